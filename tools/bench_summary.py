"""Print the headline numbers of a bench.py JSON line (profiling aid)."""
import json
import sys

j = json.load(open(sys.argv[1]))
print(f"rays/s {j['value']:.4g}  ms/step {j['ms_per_step']:.1f}  e2e {j['e2e']['value']:.4g}  accuracy {j.get('accuracy')}")
for k in j["kernels"]:
    print("  ", k)
for r in j["rooflines"]:
    print(f"   {r['kernel']}: frac {r['frac']:.3f} share {r['share_of_step']:.3f}")
