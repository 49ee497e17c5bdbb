"""paillier_halo2_b200 — B200-native batched Paillier encrypt / add / tally and PaillierChip witnesses.

Only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of include/paillier_b200.h), `api.py`
(host-side mirror of the reference's paillier_enc_native / paillier_add_native / PaillierChip for this
path), `workload.py` (the seeded synthetic inputs of SURVEY.md §8d) and `build.py`.
"""
from .api import PaillierKey, Pb200Error, ints_to_words, words_to_ints, witness_digest  # noqa: F401

__all__ = ["PaillierKey", "Pb200Error", "ints_to_words", "words_to_ints", "witness_digest"]
