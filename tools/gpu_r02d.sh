# round 2, call D: whole GPU suite, then the bench line (scheduler-run sweep slice, fp32 mode, HBM kernels, long roofline pass)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02d_tests.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r02d_tests.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r02d_bench.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02d_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"])
r = d["roofline"]; print("roofline", r["achieved"], r["frac"], r.get("frac_of_burst_peak"), r.get("pass_seconds"), r.get("pass_clocks"))
for s in r["by_shape"]: print(s)
print("fp32", d.get("fp32_mode"))
print("sweep", {k: v for k, v in d.get("sweep", {}).items() if k not in ("steady_state_replica", "what")})
print("steady", {k: v for k, v in d.get("sweep", {}).get("steady_state_replica", {}).items() if k.startswith(("sec", "cond"))})
print("vit", d.get("vit_b16", {}).get("value"))
for k in d.get("roofline_hbm", {}).get("kernels", []): print(k)
print("hbm err", d.get("roofline_hbm", {}).get("error"))
print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("sample_batch"))
PY
