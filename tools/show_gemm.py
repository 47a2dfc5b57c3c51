import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "ERR", e)
        continue
    print(f)
    for k, v in d["shapes"].items():
        print(f"  {k:20s} {v['ms']*1e3:8.1f}us {v['tflops']:7.0f}TF clk={v['sm_mhz']} W={v['watts']} "
              f"frac={v.get('frac_of_clock_peak', 0):.3f} cublas={v.get('cublas_plain_tflops', 0):.0f}@{v.get('cublas_sm_mhz')}")
