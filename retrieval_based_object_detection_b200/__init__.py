"""B200-native retrieval hot path for dmweapon/Retrieval_based_object_detection.

Public surface:
  * ``Gallery``            -- device-resident collection vectors (K1 upsert, K2 segment mean, K3 search)
  * ``merge_topk``         -- K4 merge of per-GPU top-k lists
  * ``ShardedGallery``     -- row-sharded multi-GPU search and delegate build over torch.distributed / NCCL
  * the drop-in ``qdrant_client`` package at the repo root is built on these.

The arithmetic lives in librbod.so (hand-written sm_100a CUDA behind the C ABI of include/rbod.h).
Importing this package does not load the library; the first Gallery does, and fails loudly if the
extension has not been built or no B200 is present.
"""
from .gallery import Gallery, SearchResult, merge_topk, merge_topk_packed, l2norm_pack, segment_finish  # noqa: F401
from .sharded import ShardedGallery, shard_range  # noqa: F401

__all__ = ["Gallery", "SearchResult", "merge_topk", "merge_topk_packed", "l2norm_pack", "segment_finish", "ShardedGallery", "shard_range"]
__version__ = "0.1.0"
