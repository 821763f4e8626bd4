"""Runs the reference's four hot-path scripts UNMODIFIED against the drop-in ``qdrant_client`` / ``clip`` packages,
driven through their interactive prompts by piped stdin, on a synthetic image tree (tests/ref_chain.py).

Builder-container test (``-m "not gpu"``): the scripts live in /root/reference, which does not exist on the GPU
box, and this container has no GPU, so vector arithmetic is answered by the Gallery test double
(tests/fakes/fake_gallery.py = the oracle) installed by tests/fakes/run_script.py.  What is proven here is the
drop-in boundary: every call the scripts make, with their argument spellings and access patterns, and the chain of
separate processes sharing one on-disk store.  tests/test_gpu_reference_scripts.py runs the same chain on a B200
against librbod.so when the scripts are shipped as job inputs; tests/test_gpu_shim.py replays their call sequences.
"""
import os
import sys

import pytest

import ref_chain
from oracle import ref_loader as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
RUNNER = os.path.join(ROOT, "tests", "fakes", "run_script.py")

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present (builder container only)")


def test_manager_31_32_33_run_unchanged(tmp_path):
    work, store = tmp_path / "work", tmp_path / "store"
    work.mkdir()
    ref_chain.make_images(work)
    rows, n_points = ref_chain.run_chain(REF, work, store, runner=RUNNER, cpu_only=True)

    # ---- check the stored state and the scripts' numbers against the reference's own functions, from a fresh "process"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fakes import fake_gallery

    import retrieval_based_object_detection_b200.gallery as gallery

    saved = gallery.Gallery
    gallery.Gallery = fake_gallery.FakeGallery
    os.environ["RBOD_STORE_DIR"] = str(store)
    try:
        import qdrant_client as qc

        ref32 = R.delegate_module()
        fns = (("average", ref32.compute_average), ("centroid", ref32.compute_centroid),
               ("weighted", ref32.compute_weighted_average), ("medoid", ref32.compute_medoid))
        ref_chain.check_store(qc.QdrantClient(host="localhost", port=6333), rows, n_points, fns,
                              ref32.generate_delegate_id, R.cosine_similarity())
        qc._close_all()
    finally:
        gallery.Gallery = saved
        os.environ.pop("RBOD_STORE_DIR", None)
