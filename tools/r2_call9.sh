#!/usr/bin/env bash
# round 2, GPU call 9 (2 GPUs): flag-in-data peer exchange: parity + TP-2 bench, CLI --tp 2 against --tp 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -q -x > gpurun_out/c9_tp.log 2>&1; echo "tp rc=$?"; tail -6 gpurun_out/c9_tp.log
for nccl in 0 1; do
NT_B200_TP_NCCL=$nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/c9_bench_tp2_nccl$nccl.json 2> gpurun_out/c9_bench_tp2_nccl$nccl.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c9_bench_tp2_nccl$nccl.json").read().strip().splitlines()[-1]); print("nccl=$nccl", d["value"], d["ms_per_step"], d["e2e"]["value"], d["tp"], d["path"])
except Exception as e: print("nccl=$nccl failed", e); print(open("gpurun_out/c9_bench_tp2_nccl$nccl.err").read()[-2000:])
PY
done
python - <<'PY' > gpurun_out/c9_cli.log 2>&1
import subprocess, sys
sys.path.insert(0, ".")
from ntransformer_b200.gguf_write import synthetic_tensors_np, write_gguf
from ntransformer_b200.model_spec import LlamaConfig
cfg = LlamaConfig(vocab_size=2048, hidden_size=2048, intermediate_size=4096, n_layers=3, n_heads=16, n_kv_heads=8, head_dim=128, max_seq_len=128, bos_token_id=1, eos_token_id=2)
write_gguf("/tmp/cli_tp.gguf", cfg, synthetic_tensors_np(cfg, "Q4_K_M", seed=21))
outs = []
for tp in (1, 2):
    r = subprocess.run(["./ntransformer_b200/ntransformer", "-m", "/tmp/cli_tp.gguf", "-p", "hello world", "-n", "24", "-t", "0", "--repeat-penalty", "1.0", "-c", "128", "--tp", str(tp)], capture_output=True, text=True, timeout=300)
    print("tp", tp, "rc", r.returncode, repr(r.stdout[-400:]), r.stderr[-600:])
    outs.append(r.stdout)
print("CLI_TP_MATCH", outs[0] == outs[1] and len(outs[0]) > 0)
PY
tail -12 gpurun_out/c9_cli.log
