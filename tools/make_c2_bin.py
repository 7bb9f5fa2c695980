"""Write the C2 workload (x, y, err, start) as a flat binary for tools/variant_bench."""
import os, struct, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datasets, oracle_lib as oracle
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
model, x, y, err = datasets.dataset(name)
start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
with open(os.path.join(ROOT, "tools", name + ".bin"), "wb") as f:
    f.write(struct.pack("ii", len(x), len(start)))
    for a in (x, y, err, start):
        f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
print("wrote", name, len(x), start)
