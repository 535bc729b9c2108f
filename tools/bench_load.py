"""Loader throughput: writes the workload's seeded GGUF to /dev/shm, then times our pipelined loader
(nt_model_load_gguf: reader threads -> pinned staging -> cudaMemcpyAsync) and the reference's own loader (oracle/_ref:
Transformer::load, one synchronous cudaMemcpy per tensor from the mmap) on the same file.  Prints one JSON line.

    python tools/bench_load.py --workload llama3-8b-q4_k_m-decode [--threads 8]"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="llama3-8b-q4_k_m-decode")
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    os.environ["NT_LOAD_THREADS"] = str(args.threads)
    import torch
    import bench as B
    from ntransformer_b200.engine import Model
    from ntransformer_b200.gguf_write import write_gguf_streaming
    from ntransformer_b200.model_spec import tensor_table
    from ntransformer_b200.synth import random_blocks_cuda

    shape, mix, _, max_seq = B.WORKLOADS[args.workload]
    cfg = B.shape_cfg(shape, max_seq)
    path = f"/dev/shm/nt_load_{os.getpid()}.gguf"

    def gen():
        for idx, (name, dt, rows, cols) in enumerate(tensor_table(cfg, mix)):
            if name.endswith("norm.weight"):
                t = 1.0 + 0.1 * torch.randn(cols, device="cuda", dtype=torch.float32)
            else:
                t = random_blocks_cuda(dt, rows, cols, 1234 + idx * 16)
            yield name, t.cpu().numpy(), dt, rows, cols
    t0 = time.time()
    write_gguf_streaming(path, cfg, gen(), tensor_table(cfg, mix))
    gen_s = time.time() - t0
    size = os.path.getsize(path)
    out = {"workload": args.workload, "file_gb": round(size / 1e9, 2), "write_s": round(gen_s, 1), "reader_threads": args.threads}
    try:
        log = f"/tmp/nt_load_{os.getpid()}.log"
        devnull, saved = os.open(log, os.O_WRONLY | os.O_CREAT | os.O_TRUNC), os.dup(2)
        os.dup2(devnull, 2)
        try:
            for rep in range(2):                       # second pass: page cache certainly warm for both
                t0 = time.time()
                m = Model.load(path, max_seq)
                torch.cuda.synchronize()
                ours = time.time() - t0
                inner = m.load_seconds
                m.close()
            ref_s = None
            so = ROOT / "oracle" / "_ref" / "libnt_ref.so"
            if so.exists() and not args.no_reference:
                ref = C.CDLL(str(so))
                ref.ref_model_load.restype = C.c_void_p
                ref.ref_model_load.argtypes = [C.c_char_p, C.c_int]
                ref.ref_model_free.argtypes = [C.c_void_p]
                t0 = time.time()
                h = ref.ref_model_load(path.encode(), max_seq)
                torch.cuda.synchronize()
                ref_s = time.time() - t0
                ref.ref_model_free(h)
        finally:
            os.dup2(saved, 2)
            os.close(devnull)
    finally:
        os.unlink(path)
    try:
        out["ours_copy_phase"] = [l.strip() for l in open(log, errors="replace") if "reader threads" in l][-1]
        os.unlink(log)
    except Exception:
        pass
    out.update({"ours_load_s": round(ours, 2), "ours_load_s_inner": round(inner, 2), "ours_gb_per_s": round(size / 1e9 / ours, 2),
                "reference_load_s": None if ref_s is None else round(ref_s, 2),
                "reference_gb_per_s": None if ref_s is None else round(size / 1e9 / ref_s, 2)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
