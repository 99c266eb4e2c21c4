"""
Regenerates tests/golden/* from the read-only reference checkout (run in the BUILD container
only; /root/reference does not exist on the GPU box, which is why the outputs are committed).

  ref_test_{ei,poi,cb}.npy  <- byte copies of the reference's own golden fixtures
                               /root/reference/test/test_data/test_{ei,poi,cb}.npy
                               (consumed by test/test_boptim.py:42-58)
  notebook_kat_ei.json      <- "Final parameter values" lines stored in the output of
                               examples/notebooks/GP_based_exploration_exploitation.ipynb cell 13
                               (boptimizer, EI, np.random.seed(42) seeds, CPU fp64)
  oracle_*.npz              <- outputs of oracle/gp_oracle.py on seeded inputs (make_oracle_vectors.py, same directory)
"""
import hashlib
import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def copy_reference_goldens():
    out = {}
    for name in ("ei", "poi", "cb"):
        src = os.path.join(REF, "test", "test_data", f"test_{name}.npy")
        dst = os.path.join(HERE, f"ref_test_{name}.npy")
        shutil.copyfile(src, dst)
        out[f"ref_test_{name}.npy"] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    return out


def notebook_kat():
    nb = json.load(open(os.path.join(REF, "examples", "notebooks",
                                     "GP_based_exploration_exploitation.ipynb")))
    cell = nb["cells"][13]
    txt = "".join("".join(o.get("text", "")) for o in cell["outputs"])
    rows = re.findall(r"amp: ([\d.e-]+), lengthscale: \[\s*([\d.e-]+)\s+([\d.e-]+)\], noise: ([\d.e-]+)", txt)
    kat = [{"amp": float(a), "lengthscale": [float(b), float(c)], "noise": float(d)} for a, b, c, d in rows]
    json.dump({"source": "examples/notebooks/GP_based_exploration_exploitation.ipynb cell 13 stored output",
               "setup": "25x25 three-Gaussian trial_func, np.random.seed(42) randint(0,25,(5,2)) seeds, "
                        "boptimizer(acquisition_function='ei', exploration_steps=50, use_gpu=False)",
               "trainings": kat}, open(os.path.join(HERE, "notebook_kat_ei.json"), "w"), indent=1)
    return len(kat)


if __name__ == "__main__":
    print(copy_reference_goldens())
    print("notebook KAT trainings:", notebook_kat())
    if "--oracle" in sys.argv:
        sys.path.insert(0, HERE)
        from make_oracle_vectors import main
        main()
