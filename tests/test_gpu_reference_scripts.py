"""The reference's four hot-path scripts, UNMODIFIED, on the product path: a B200 and librbod.so underneath the drop-in
``qdrant_client`` (no test double anywhere), each script its own process, all sharing one on-disk store.

The scripts are not part of this repository and /root/reference does not exist on a GPU box, so they travel as job
inputs: ``tools/stage_reference_scripts.sh`` copies them into the git-ignored ``gpurun_in/reference_scripts/`` of the
builder container, which gpurun ships with the snapshot.  Without them (the driver's own GPU run) the test is skipped;
tests/test_gpu_shim.py then still replays the scripts' call sequences against the library.
"""
import os

import pytest

import ref_chain
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RBOD_REFERENCE_SCRIPTS") or os.path.join(ROOT, "gpurun_in", "reference_scripts")
NEEDED = ("util/qdrant_manager.py", "31_clip_embedding_and_save_vector.py", "32_create_delegate_vector.py",
          "33_run_all_experiments.py")


@pytest.mark.skipif(not all(os.path.isfile(os.path.join(REF, s)) for s in NEEDED),
                    reason="the reference scripts were not shipped as job inputs (tools/stage_reference_scripts.sh)")
def test_unmodified_scripts_run_on_the_b200_library(tmp_path, store_dir):
    work = tmp_path / "work"
    work.mkdir()
    ref_chain.make_images(work)
    log_path = os.path.join(ROOT, "gpurun_out", "reference_scripts_on_b200.log")
    os.makedirs(os.path.dirname(log_path), exist_ok=True)
    with open(log_path, "w", encoding="utf-8") as log:
        rows, n_points = ref_chain.run_chain(REF, work, store_dir, runner=None, cpu_only=False, log=log)
        # a fresh client in this process (the real Gallery): what the scripts stored and computed, against the oracle's
        # restatements of the reference functions (32_...py:9-26, 33_...py:76-77; pinned bit-for-bit to the originals by
        # tests/test_oracle_golden.py)
        import qdrant_client as qc

        fns = (("average", O.compute_average), ("centroid", O.compute_centroid),
               ("weighted", O.compute_weighted_average), ("medoid", O.compute_medoid))
        client = qc.QdrantClient(host="localhost", port=6333)
        ref_chain.check_store(client, rows, n_points, fns, None, O.cosine_similarity, ulp_tol=1)
        col = client._root.get("thesis")
        assert type(col.gallery).__module__.endswith("gallery") and type(col.gallery).__name__ == "Gallery"
        info = col.gallery.info()
        log.write(f"\n===== check =====\n{len(rows)} CSV rows, {n_points} image points + 16 delegates verified against "
                  f"the oracle through librbod.so (device {info['device']}, rows {info['rows']}, dim {info['dim']})\n")
