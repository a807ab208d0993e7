"""zkb200 -- B200-native accelerator for the groth16::prove() hot path of republicprotocol/zksnark-rs.

The product is ``libzkb200.so`` (hand-written sm_100a CUDA behind the C ABI in ``include/zkb200.h``).
This package is the thin host-side mirror of the reference's interface for that path
(``groth16::{setup, prove, verify}``, ``QAP``, ``SigmaG1/SigmaG2``, ``Proof``; src/groth16/mod.rs) on top of
that library via ctypes.  There is no CPU fallback: importing works anywhere, but every call that
computes raises ``ZkbError`` unless the CUDA library is built and a B200 is present.
"""

from .groth16 import (CRS, QAP, Bases, Context, Proof, ZkbError, fr_limbs, horner_qap_rows, lib_path,
                      load_library, msm, ntt, prove, prove_batch, prove_partial, prove_combine, prove_combine_batch, qap_h, setup, verify, verify_batch, pairing,
                      Comm, setup_shard, crs_upload_shard, prove_shard, prove_shard_batch, prove_shard_enqueue, prove_shard_collect, ntt_shard,
                      WitnessPlan, weights, witness_generate_raw, witness_generate_dev, layered_qap_rows)

__all__ = ["CRS", "QAP", "Bases", "Context", "Proof", "ZkbError", "fr_limbs", "horner_qap_rows", "lib_path",
           "load_library", "msm", "ntt", "prove", "prove_batch", "prove_partial", "prove_combine", "prove_combine_batch", "qap_h", "setup", "verify", "verify_batch", "pairing",
           "Comm", "setup_shard", "crs_upload_shard", "prove_shard", "prove_shard_batch", "prove_shard_enqueue", "prove_shard_collect", "ntt_shard",
           "WitnessPlan", "weights", "witness_generate_raw", "witness_generate_dev", "layered_qap_rows"]
