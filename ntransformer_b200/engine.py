"""Host-side mirror of the reference's nt::Transformer surface (src/model/transformer.h:31-61) over the native
engine's C-ABI (include/nt_b200_engine.h): load / forward(tokens, start_pos) -> logits.

torch is used only to own device memory for synthetic weights and for torch.distributed plumbing (exchange
of the NCCL id under tensor parallelism)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._engine_sigs import ModelConfigC
from ._lib import lib
from .dtypes import DType
from .model_spec import LlamaConfig, tensor_table
from .tp import split_kind


class Model:
    def __init__(self, handle, cfg: LlamaConfig, keep=None, tp_rank=0, tp_size=1):
        self._h = handle
        self.cfg = cfg
        self._keep = keep          # device tensors backing borrowed weights
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self._logits_host = np.empty(cfg.vocab_size, dtype=np.float32)

    # ---- construction ----
    @classmethod
    def load(cls, gguf_path: str, max_context: int = 4096, tp_rank: int = 0, tp_size: int = 1) -> "Model":
        h = lib().nt_model_load_gguf(str(gguf_path).encode(), max_context, tp_rank, tp_size)
        if not h:
            raise RuntimeError(f"failed to load {gguf_path}")
        c = ModelConfigC()
        lib().nt_model_get_config(h, C.byref(c))
        cfg = LlamaConfig(**{n: getattr(c, n) for n, _ in ModelConfigC._fields_})
        return cls(h, cfg, tp_rank=tp_rank, tp_size=tp_size)

    @classmethod
    def from_device_tensors(cls, cfg: LlamaConfig, tensors: dict, tp_rank: int = 0, tp_size: int = 1) -> "Model":
        """tensors: gguf name -> (torch uint8/float32 CUDA tensor, DType); shapes already sharded for tp_rank."""
        c = ModelConfigC(**cfg.dict())
        h = lib().nt_model_create(C.byref(c), tp_rank, tp_size)
        for name, spec in tensors.items():
            t, dt = spec[0], spec[1]
            pitch = spec[2] if len(spec) > 2 else 0          # bytes between rows (0 = dense GGUF rows)
            rc = lib().nt_model_set_tensor(h, name.encode(), t.data_ptr(), int(dt), pitch)
            if rc != 0:
                raise ValueError(f"engine rejected tensor {name}")
        if lib().nt_model_finalize(h) != 0:
            lib().nt_model_free(h)
            raise RuntimeError("model is missing tensors")
        return cls(h, cfg, keep=tensors, tp_rank=tp_rank, tp_size=tp_size)

    @classmethod
    def synthetic(cls, cfg: LlamaConfig, mix: str, seed: int = 1234, tp_rank: int = 0, tp_size: int = 1, device="cuda") -> "Model":
        """Random valid GGUF blocks generated directly on the GPU (no checkpoint or file involved)."""
        import torch

        from .synth import random_blocks_cuda

        tensors = {}
        for idx, (name, dt, rows, cols) in enumerate(tensor_table(cfg, mix, tp_rank, tp_size)):
            # replicated tensors share a seed across ranks, shards get rank-specific seeds
            replicated = name.endswith("norm.weight") or name == "token_embd.weight"
            s = seed + idx * 16 + (0 if replicated else tp_rank)
            if name.endswith("norm.weight"):
                g = torch.Generator(device=device)
                g.manual_seed(s)
                t = 1.0 + 0.1 * torch.randn(cols, generator=g, device=device, dtype=torch.float32)
            else:
                t = random_blocks_cuda(dt, max(rows, 1), cols, s, device=device)
            t = t.contiguous()
            pad_all = os.environ.get("NT_B200_SYNTH_PAD") == "1" and name != "token_embd.weight"      # tools/prof_decode.py --shard-of
            if t.dim() == 2 and t.dtype == torch.uint8 and t.shape[1] % 16 and (pad_all or (tp_size > 1 and split_kind(name) == "cols")):
                # rows whose byte length is not a multiple of 16 (column shards of Q6_K, narrow models): pad the row pitch so rows
                # stay TMA-aligned (what the GGUF loader does for column shards too)
                pitch = (t.shape[1] + 15) // 16 * 16
                padded = torch.zeros((t.shape[0], pitch), dtype=torch.uint8, device=device)
                padded[:, : t.shape[1]] = t
                tensors[name] = (padded, dt, pitch)
            else:
                tensors[name] = (t, dt)
        torch.cuda.synchronize()
        return cls.from_device_tensors(cfg, tensors, tp_rank, tp_size)

    # ---- tensor parallel ----
    def init_tp(self):
        """Collective: exchanges the NCCL unique id over torch.distributed and builds the communicator."""
        import torch
        import torch.distributed as dist

        buf = (C.c_ubyte * 128)()
        if self.tp_rank == 0 and lib().nt_tp_unique_id(buf) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        idb = (C.c_ubyte * 128)(*t.cpu().tolist())
        rc = lib().nt_tp_init(self._h, idb, self.tp_rank, self.tp_size)
        if rc != 0:
            raise RuntimeError(f"nt_tp_init failed ({rc})")

    # ---- inference ----
    def forward(self, tokens, start_pos: int, want_logits: bool = True):
        """Runs the tokens at positions start_pos.. and returns HOST logits [vocab] of the last one."""
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        out = self._logits_host.ctypes.data_as(C.c_void_p) if want_logits else None
        rc = lib().nt_model_forward(self._h, toks.ctypes.data_as(C.c_void_p), len(toks), start_pos, out)
        assert rc == 0
        return self._logits_host if want_logits else None

    def forward_async(self, tokens, start_pos: int):
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        lib().nt_model_forward_async(self._h, toks.ctypes.data_as(C.c_void_p), len(toks), start_pos)

    def sync(self):
        rc = lib().nt_model_sync(self._h)
        if rc != 0:
            raise RuntimeError(f"CUDA error {rc} on the model stream")

    def argmax(self) -> int:
        return lib().nt_model_argmax(self._h)

    def clear_kv(self):
        lib().nt_model_clear_kv(self._h)

    def use_graph(self, on: bool):
        lib().nt_model_use_graph(self._h, int(on))

    def set_prefill_min_tokens(self, n: int):
        """Prompts of >= n tokens take the batched tensor-core prefill (F16 models); 0 = per-token replay only."""
        lib().nt_model_set_prefill_min_tokens(self._h, int(n))

    def use_megakernel(self, on: bool = True):
        """Opt-in: decode each token as one persistent kernel (csrc/engine/decode_mega.h). Call before the first forward."""
        lib().nt_model_use_megakernel(self._h, int(on))

    @property
    def load_seconds(self) -> float:
        return float(lib().nt_model_load_seconds(self._h))

    @property
    def tp_exchange(self) -> str:
        return {0: "none", 1: "ncclAllReduce", 2: "nvlink peer-memory exchange (GEMV epilogue)"}[int(lib().nt_model_tp_exchange(self._h))]

    @property
    def megakernel_active(self) -> bool:
        return bool(lib().nt_model_megakernel_active(self._h))

    def megakernel_plan(self):
        """Phase kinds of the persistent kernel's per-token program ([] when it is not active)."""
        n = lib().nt_model_megakernel_plan(self._h, None, 0)
        if n <= 0:
            return []
        buf = (C.c_int * n)()
        lib().nt_model_megakernel_plan(self._h, buf, n)
        return list(buf)

    def megakernel_trace(self, on: bool = True):
        """Record the persistent kernel's phase timeline (call after it is active, i.e. after one decoded token)."""
        lib().nt_model_megakernel_trace(self._h, int(on))

    def megakernel_trace_read(self):
        """(ticks [4 CTAs, phases, (start, work done, barrier passed)] in SM clock ticks, ns_per_tick [4]) of the most recent launch."""
        n = lib().nt_model_megakernel_trace_read(self._h, None, 0)
        if n <= 0:
            return np.zeros((0, 0, 3), dtype=np.uint64), np.zeros(0)
        out = np.zeros(n, dtype=np.uint64)
        lib().nt_model_megakernel_trace_read(self._h, out.ctypes.data_as(C.c_void_p), n)
        per = out.reshape(4, -1)
        cal = per[:, -4:].astype(np.float64)                     # clock start/end, globaltimer start/end
        ns_per_tick = (cal[:, 3] - cal[:, 2]) / np.maximum(cal[:, 1] - cal[:, 0], 1.0)
        return per[:, :-4].reshape(4, -1, 3), ns_per_tick

    def debug_read(self, name: str):
        """Host copy of one of the persistent kernel's working vectors (hid0, hid1, q, attn, act, slots)."""
        n = lib().nt_model_debug_read(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float32)
        lib().nt_model_debug_read(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), n)
        return out

    @property
    def stream(self) -> int:
        return lib().nt_model_stream(self._h)

    @property
    def logits_device_ptr(self) -> int:
        return lib().nt_model_logits_device(self._h)

    def bytes_per_token(self, ctx: int) -> int:
        return int(lib().nt_model_bytes_per_token(self._h, ctx))

    def close(self):
        if self._h:
            lib().nt_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """include/ntransformer.h through ctypes (the reference declares this C API but never implements it)."""

    def __init__(self):
        self._h = lib().nt_engine_create()

    def load(self, path: str) -> bool:
        return lib().nt_engine_load(self._h, str(path).encode()) == 0

    def generate(self, prompt: str, max_tokens=32, temperature=0.0, top_k=40, top_p=0.9) -> str:
        p = lib().nt_engine_generate(self._h, prompt.encode(), max_tokens, temperature, top_k, top_p)
        if not p:
            raise RuntimeError("generate failed")
        s = C.cast(p, C.c_char_p).value.decode(errors="replace")
        lib().nt_free(p)
        return s

    @property
    def vocab_size(self):
        return lib().nt_engine_vocab_size(self._h)

    @property
    def n_layers(self):
        return lib().nt_engine_n_layers(self._h)

    @property
    def hidden_size(self):
        return lib().nt_engine_hidden_size(self._h)

    def close(self):
        if self._h:
            lib().nt_engine_destroy(self._h)
            self._h = None
