mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "sharded_over_two" > gpurun_out/r02m_pytest_2gpu.log 2>&1; tail -n 3 gpurun_out/r02m_pytest_2gpu.log
(time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline) > gpurun_out/r02m_bench_2gpu.json 2> gpurun_out/r02m_bench_2gpu.err; tail -n 4 gpurun_out/r02m_bench_2gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/r02m_bench_2gpu.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], d["ms_steps"], "e2e", d["e2e"]["value"], d["e2e"]["ms_steps"])
        for k, v in d.get("other_workloads", {}).items():
            print(" ", k, json.dumps(v)[:700])
PY
