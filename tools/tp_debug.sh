#!/bin/bash
# quick 2-GPU tensor-parallel bring-up check (bounded time)
set -x
python - <<'PY'
import sys; sys.path.insert(0,'.')
from ntransformer_b200.gguf_write import synthetic_tensors_np, write_gguf
from ntransformer_b200.model_spec import LlamaConfig
CFG = LlamaConfig(vocab_size=2048, hidden_size=2048, intermediate_size=4096, n_layers=3, n_heads=16, n_kv_heads=4, head_dim=128, max_seq_len=128, bos_token_id=1, eos_token_id=2)
write_gguf('/tmp/tp.gguf', CFG, synthetic_tensors_np(CFG, 'Q4_K_M', seed=21))
PY
NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/tp_worker.py /tmp/tp.gguf /tmp/tp_out.npz 128 2>&1 | grep -vE "^(  Loaded|===|GGUF|Layers|Max seq|RoPE|BOS|File size|Tensor data|Tensors:|Vocab|Architecture|Name:|Loading|Note)" | tail -25
echo "graph run rc=$?"
if [ ! -f /tmp/tp_out.npz ]; then
NT_B200_NO_GRAPH=1 NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/tp_worker.py /tmp/tp.gguf /tmp/tp_out.npz 128 2>&1 | grep -vE "^(  Loaded|===|GGUF|Layers|Max seq|RoPE|BOS|File size|Tensor data|Tensors:|Vocab|Architecture|Name:|Loading|Note)" | tail -25
echo "eager run done"
fi
ls -la /tmp/tp_out.npz
