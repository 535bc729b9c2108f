"""ctypes/numpy wrapper of the CPU oracle (oracle/libnt_oracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/nt_oracle.h.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libnt_oracle.so"
_lib = None

F32, F16, Q8_0, Q4_0, Q4_K, Q6_K, Q5_K = 0, 1, 2, 3, 4, 5, 6


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("vocab_size", "hidden_size", "intermediate_size", "n_layers", "n_heads",
                                       "n_kv_heads", "head_dim", "max_seq_len")] + [("norm_eps", C.c_float), ("rope_theta", C.c_float)]


class Layer(C.Structure):
    _fields_ = [("attn_norm", C.c_void_p), ("ffn_norm", C.c_void_p)] + \
               [(n, C.c_void_p) for n in ("wq", "wk", "wv", "wo", "w_gate", "w_up", "w_down")] + \
               [(n, C.c_int) for n in ("dt_q", "dt_k", "dt_v", "dt_o", "dt_gate", "dt_up", "dt_down")]


def build() -> Path:
    src = HERE / "nt_oracle.c"
    if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, (HERE / "nt_oracle.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(HERE), "libnt_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB_PATH))
        L = _lib
        L.nto_fp16_to_fp32.restype = C.c_float
        L.nto_fp16_to_fp32.argtypes = [C.c_uint16]
        L.nto_fp32_to_fp16.restype = C.c_uint16
        L.nto_fp32_to_fp16.argtypes = [C.c_float]
        L.nto_row_bytes.restype = C.c_size_t
        L.nto_row_bytes.argtypes = [C.c_int, C.c_int64]
        L.nto_dtype_size.restype = C.c_size_t
        L.nto_dtype_size.argtypes = [C.c_int]
        L.nto_dtype_block_size.restype = C.c_size_t
        L.nto_dtype_block_size.argtypes = [C.c_int]
        L.nto_model_create.restype = C.c_void_p
        L.nto_model_create.argtypes = [C.POINTER(Config)]
        L.nto_model_destroy.argtypes = [C.c_void_p]
        L.nto_model_set_globals.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.nto_model_set_layer.argtypes = [C.c_void_p, C.c_int, C.POINTER(Layer)]
        L.nto_model_forward.restype = C.c_int
        L.nto_model_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.nto_num_threads.restype = C.c_int
    return _lib


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def row_bytes(dt: int, n: int) -> int:
    return lib().nto_row_bytes(dt, n)


def dequant_rows(dt: int, raw: np.ndarray, rows: int, cols: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw).view(np.uint8).reshape(rows, -1)
    out = np.empty((rows, cols), dtype=np.float32)
    L = lib()
    for r in range(rows):
        L.nto_dequant_row(C.c_int(dt), _ptr(raw[r]), C.c_int64(cols), _ptr(out[r]))
    return out


def gemv(W: np.ndarray, x: np.ndarray, out_features: int, in_features: int, dt: int) -> np.ndarray:
    W = np.ascontiguousarray(W)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros(out_features, dtype=np.float32)
    lib().nto_gemv(_ptr(y), _ptr(W), _ptr(x), C.c_int(out_features), C.c_int(in_features), C.c_int(dt))
    return y


def rmsnorm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, hidden = (1, x.shape[0]) if x.ndim == 1 else x.shape
    y = np.empty_like(x)
    lib().nto_rmsnorm(_ptr(y), _ptr(x), _ptr(np.ascontiguousarray(w, dtype=np.float32)), C.c_int(rows), C.c_int(hidden), C.c_float(eps))
    return y


def rope(q, k, positions, n_heads, n_kv_heads, head_dim, theta, freq_scale=1.0, interleaved=False):
    q = np.array(q, dtype=np.float32, copy=True)
    k = np.array(k, dtype=np.float32, copy=True)
    pos = np.ascontiguousarray(positions, dtype=np.int32)
    lib().nto_rope(_ptr(q), _ptr(k), _ptr(pos), C.c_int(len(pos)), C.c_int(n_heads), C.c_int(n_kv_heads), C.c_int(head_dim),
                   C.c_float(theta), C.c_float(freq_scale), C.c_int(int(interleaved)))
    return q, k


def copy_to_kv_cache(kc, vc, k, v, seq_len, n_kv, hd, start_pos, max_seq):
    lib().nto_copy_to_kv_cache(_ptr(kc), _ptr(vc), _ptr(np.ascontiguousarray(k, dtype=np.float32)),
                               _ptr(np.ascontiguousarray(v, dtype=np.float32)), C.c_int(seq_len), C.c_int(n_kv), C.c_int(hd),
                               C.c_int(start_pos), C.c_int(max_seq))


def attention_decode(q, kc, vc, seq_len, n_heads, n_kv, hd, max_seq, scale):
    out = np.empty((n_heads, hd), dtype=np.float32)
    lib().nto_attention_decode(_ptr(out), _ptr(np.ascontiguousarray(q, dtype=np.float32)), _ptr(kc), _ptr(vc), C.c_int(seq_len),
                               C.c_int(n_heads), C.c_int(n_kv), C.c_int(hd), C.c_int(max_seq), C.c_float(scale))
    return out


def attention_prefill(Q, kc, vc, seq_len, start_pos, n_heads, n_kv, hd, max_seq, scale):
    out = np.empty((seq_len, n_heads, hd), dtype=np.float32)
    lib().nto_attention_prefill(_ptr(out), _ptr(np.ascontiguousarray(Q, dtype=np.float32)), _ptr(kc), _ptr(vc), C.c_int(seq_len),
                                C.c_int(start_pos), C.c_int(n_heads), C.c_int(n_kv), C.c_int(hd), C.c_int(max_seq), C.c_float(scale))
    return out


def silu_mul(gate, up):
    gate = np.ascontiguousarray(gate, dtype=np.float32)
    out = np.empty_like(gate)
    lib().nto_silu_mul(_ptr(out), _ptr(gate), _ptr(np.ascontiguousarray(up, dtype=np.float32)), C.c_int(gate.size))
    return out


def fp16_bits_to_f32(h: np.ndarray) -> np.ndarray:
    L = lib()
    return np.array([L.nto_fp16_to_fp32(int(v)) for v in np.asarray(h, dtype=np.uint16).ravel()], dtype=np.float32).reshape(np.shape(h))


def f32_to_fp16_bits(f: np.ndarray) -> np.ndarray:
    L = lib()
    return np.array([L.nto_fp32_to_fp16(float(v)) for v in np.asarray(f, dtype=np.float32).ravel()], dtype=np.uint16).reshape(np.shape(f))


class Model:
    """Whole-model CPU forward (reference src/model/transformer.cpp:604-669) over host numpy weights.

    `weights` maps GGUF tensor names to (uint8/float array, dtype id)."""

    def __init__(self, cfg: dict, weights: dict):
        self.cfg = Config(**{k: cfg[k] for k, _ in Config._fields_})
        self._keep = weights
        self.h = lib().nto_model_create(C.byref(self.cfg))
        te, dte = weights["token_embd.weight"]
        ow, dto = weights.get("output.weight", weights["token_embd.weight"])
        on, _ = weights["output_norm.weight"]
        lib().nto_model_set_globals(self.h, _ptr(te), dte, _ptr(ow), dto, _ptr(on))
        for i in range(cfg["n_layers"]):
            p = f"blk.{i}."
            L = Layer()
            L.attn_norm = _ptr(weights[p + "attn_norm.weight"][0]).value
            L.ffn_norm = _ptr(weights[p + "ffn_norm.weight"][0]).value
            for field, name in (("wq", "attn_q"), ("wk", "attn_k"), ("wv", "attn_v"), ("wo", "attn_output"),
                                ("w_gate", "ffn_gate"), ("w_up", "ffn_up"), ("w_down", "ffn_down")):
                arr, dt = weights[p + name + ".weight"]
                setattr(L, field, _ptr(arr).value)
                setattr(L, "dt_" + {"wq": "q", "wk": "k", "wv": "v", "wo": "o", "w_gate": "gate", "w_up": "up", "w_down": "down"}[field], dt)
            lib().nto_model_set_layer(self.h, i, C.byref(L))

    def forward(self, tokens, start_pos: int, n_layers_run: int = -1) -> np.ndarray:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        logits = np.empty(self.cfg.vocab_size, dtype=np.float32)
        rc = lib().nto_model_forward(self.h, _ptr(toks), len(toks), start_pos, _ptr(logits), n_layers_run)
        assert rc == 0
        return logits

    def close(self):
        if self.h:
            lib().nto_model_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def num_threads() -> int:
    return lib().nto_num_threads()
