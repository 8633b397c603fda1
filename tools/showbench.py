import json,sys
d=json.load(open(sys.argv[1]))
for k in ["value","ms_per_step","kernel_ms_per_step","host_ms_per_step","gpu_launches","clocks","parity","cpu_baseline"]: print(k, d.get(k))
print("e2e", d["e2e"])
print("roofline", d["roofline"])
