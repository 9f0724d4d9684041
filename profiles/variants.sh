# Tuning runs: the same bench under different builds / knobs (results never depend on them).
# usage (under gpurun): bash profiles/variants.sh "NAME|ENV=... ENV=..." ...
mkdir -p gpurun_out
for v in "$@"; do
  name="${v%%|*}"; envs="${v#*|}"
  env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-forward ${BENCH_ARGS} > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err || tail -3 gpurun_out/var_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/var_{name}.json").read().strip().splitlines()[-1])
    k = d["kernel_ms"]
    print(f"{name:14s} step {d['ms_per_step']:7.3f}  " + " ".join(f"{n[3:]}={v:.3f}" for n, v in k.items()) + f"  e2e {d['e2e']['ms_per_step']:.1f}")
except Exception as e:
    print(name, "failed", e)
PY
done
