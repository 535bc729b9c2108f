#!/usr/bin/env bash
# round 2, GPU call 21: where the per-step error spikes of the 8-layer 8B slice come from (all steps, F64 oracle beside)
mkdir -p gpurun_out
timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --layers 8 --steps 128 --oracle-steps 128 --all-errs > gpurun_out/c21_check_8l.json 2> gpurun_out/c21.err; echo "rc=$?"
NT_B200_FUSE=0 timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --layers 8 --steps 128 --all-errs > gpurun_out/c21_check_8l_unfused.json 2>> gpurun_out/c21.err; echo "rc=$?"
NT_B200_TAIL_SPLIT=0 timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --layers 8 --steps 128 --all-errs > gpurun_out/c21_check_8l_notail.json 2>> gpurun_out/c21.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("c21_check_8l","c21_check_8l_unfused","c21_check_8l_notail"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e=d["err_first_steps"]; big=[(i,x) for i,x in enumerate(e) if x>5e-4]
        print(f, "max", d["max_rel_logit_err"], "ids", d["greedy_ids_identical"], "mismatch", d["first_id_mismatch_step"], "n", len(e), "median", sorted(e)[len(e)//2], "spikes", big[:12])
        c=d.get("conditioning")
        if c: print("   ref_vs_f64 max", max(c["ref_vs_f64"]), "ours_vs_f64 max", max(c["ours_vs_f64"]), "ours spikes", [(i,x) for i,x in enumerate(c["ours_vs_f64"]) if x>5e-4][:12], "ref spikes", [(i,x) for i,x in enumerate(c["ref_vs_f64"]) if x>5e-4][:12])
    except Exception as ex: print(f, "failed", ex)
PY
