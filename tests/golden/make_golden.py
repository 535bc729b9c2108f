#!/usr/bin/env python3
"""Generates tests/golden/*.npz by IMPORTING the reference's own numpy dequantisers
(/root/reference/tools/decompose_gguf.py:219-378) in the dev container.  The reference cannot travel to
the GPU box, so the vectors are committed; re-run this script to regenerate them.

    python tests/golden/make_golden.py
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from ntransformer_b200.dtypes import DType  # noqa: E402
from ntransformer_b200.synth import random_blocks_np  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_decompose", "/root/reference/tools/decompose_gguf.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

OUT = Path(__file__).resolve().parent
ROWS, COLS = 3, 512
CASES = {
    "q6_k": (DType.Q6_K, ref.dequant_q6_k),
    "q8_0": (DType.Q8_0, ref.dequant_q8_0),
    "q4_k": (DType.Q4_K_M, ref.dequant_q4_k),
    "q5_k": (DType.Q5_K, ref.dequant_q5_k),
    "f16": (DType.F16, ref.dequant_f16),
}
for i, (name, (dt, fn)) in enumerate(CASES.items()):
    rng = np.random.default_rng(1000 + i)
    raw = random_blocks_np(dt, ROWS, COLS, rng, std=0.05)
    # make scales exercise all 6 bits / both signs
    deq = np.asarray(fn(raw.tobytes(), ROWS, COLS), dtype=np.float32)
    np.savez_compressed(OUT / f"dequant_{name}.npz", raw=raw, deq=deq, dtype=int(dt), rows=ROWS, cols=COLS)
    print(name, raw.shape, float(np.abs(deq).mean()))

# the reference's Q6_K quantiser (tools/decompose_gguf.py:389-568) round trip: quantise -> bytes
rng = np.random.default_rng(77)
w = (rng.standard_normal((2, 256)) * 0.02).astype(np.float32)
qbytes = np.frombuffer(bytes(ref.quantize_q6_k(w)), dtype=np.uint8).reshape(2, 210)
np.savez_compressed(OUT / "quantize_q6_k.npz", w=w, raw=qbytes, deq=np.asarray(ref.dequant_q6_k(qbytes.tobytes(), 2, 256), dtype=np.float32))
print("q6_k quantiser", float(np.abs(w - ref.dequant_q6_k(qbytes.tobytes(), 2, 256)).max()))
