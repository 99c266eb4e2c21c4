"""
Oracle-generated golden vectors for what NO reference test pins (Matern52 / RationalQuadratic values, 3-D
inputs, a trained fp64 trajectory, the inducing-point path, the GPyTorch parametrisation): small seeded problems,
outputs of oracle/gp_oracle.py, oracle/sparse_oracle.py and oracle/sk_oracle.py, committed as tests/golden/oracle_*.npz.  They (a) freeze the oracle itself against drift (tests/test_oracle.py) and (b) give
the GPU suite fixtures that do not depend on re-running the oracle (tests/test_gpu_parity.py).

Run from the repo root:  python tests/golden/make_oracle_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import workloads as W  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from oracle.sparse_oracle import SparseOracleGP  # noqa: E402
from oracle.sk_oracle import SKOracleGP  # noqa: E402

THETA = {"variance": 0.6, "noise": 2e-2, "scale_mixture": 1.3, "jitter": 1e-5}


def predict_case(kernel, R, ls):
    X, y = O.training_rows(O.sparse_grid(R), R)
    Xs = O.to_rows(O.full_grid(R))
    mean, sd, _ = O.predict_fixed_theta(kernel, X, y, Xs, THETA["variance"], ls, THETA["noise"], jitter=THETA["jitter"],
                                        scale_mixture=THETA["scale_mixture"])
    return {"X": X, "y": y, "Xs": Xs, "lengthscale": np.asarray(ls, dtype=np.float64), "mean": mean, "sd": sd,
            "theta": np.array([THETA["variance"], THETA["noise"], THETA["scale_mixture"], THETA["jitter"]])}


def train_case(kernel):
    R = W.dummy_blob(16, 100)
    g = O.OracleGP(O.sparse_grid(R), R, O.full_grid(R), kernel=kernel, learning_rate=0.1, iterations=20, seed=2)
    mean, sd, hp = g.run()
    return {"R": R, "mean": mean, "sd": sd, "variance": np.array(hp["variance"]), "noise": np.array(hp["noise"]),
            "lengthscale": np.array(hp["lengthscale"])}


def sparse_train_case(kernel):
    """reconstructor(sparse=True): 15 Adam steps on the VFE objective (hyper-parameters + 14 inducing inputs)."""
    R = W.dummy_blob(16, 100)
    g = SparseOracleGP(O.sparse_grid(R), R, O.full_grid(R), indpoints=14, kernel=kernel, learning_rate=0.1, iterations=15,
                       seed=2)
    mean, sd, hp = g.run()
    return {"R": R, "mean": mean, "sd": sd, "variance": np.array(hp["variance"]), "noise": np.array(hp["noise"]),
            "lengthscale": np.array(hp["lengthscale"]), "inducing_points": np.array(hp["inducing_points"]),
            "loss": np.array(g.losses)}


def sk_train_case(kernel):
    """skreconstructor(ski=False): 15 Adam steps in GPyTorch's parametrisation."""
    R = W.dummy_blob(16, 100) + 0.3
    g = SKOracleGP(O.sparse_grid(R), R, O.full_grid(R), kernel=kernel, lengthscale=[[1.0, 1.0], [10.0, 10.0]],
                   learning_rate=0.1, iterations=15)
    mean, sd, hp = g.run()
    return {"R": R, "mean": mean, "sd": sd, "noise": np.array(hp["noise"]), "lengthscale": np.array(hp["lengthscale"]),
            "loss": np.array(g.losses)}


def main():
    out = {}
    for kernel in O.KERNEL_NAMES:
        out[f"oracle_sparse_train_{kernel}.npz"] = sparse_train_case(kernel)
    for kernel in ("RBF", "Matern52"):
        out[f"oracle_sk_train_{kernel}.npz"] = sk_train_case(kernel)
    for kernel in O.KERNEL_NAMES:
        out[f"oracle_predict2d_{kernel}.npz"] = predict_case(kernel, W.dummy_blob(16, 100), [4.0, 6.0])
        out[f"oracle_predict3d_{kernel}.npz"] = predict_case(kernel, W.hyperspectral((8, 8, 6)), [2.0, 3.0, 4.0])
        out[f"oracle_train_{kernel}.npz"] = train_case(kernel)
    for name, arrays in out.items():
        np.savez_compressed(os.path.join(HERE, name), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
