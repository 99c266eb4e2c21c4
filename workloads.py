"""
Synthetic workloads of BASELINE.json `configs`.  Shared by bench.py and tests/.
Pure numpy; deterministic.  Coordinates are pixel indices (np.mgrid convention of
gprutils.get_full_grid, /root/reference/gpim/gprutils.py:136).
"""
import numpy as np

# theta used by fixed-theta timing runs: v, lengthscale per dim, noise, jitter
FIXED_THETA = {"variance": 0.05, "lengthscale": 3.0, "lengthscale_z": 8.0, "noise": 5e-3, "jitter": 1e-5}


def three_gaussians(shape):
    """Smooth field: the README's 3-Gaussian trial function (README.md:77-86) scaled to `shape`."""
    s0, s1 = shape[0] / 25.0, shape[1] / 25.0
    x, y = np.meshgrid(np.arange(shape[0]) / s0, np.arange(shape[1]) / s1, indexing="ij")

    def g(x0, y0, a, b, fwhm):
        return np.exp(-4 * np.log(2) * (a * (x - x0) ** 2 + b * (y - y0) ** 2) / fwhm ** 2)
    return g(5, 10, 1, 1, 4.5) + g(10, 8, 0.75, 1.5, 7) + g(18, 18, 1, 1.5, 10)


def spiral_mask(n, turns=15.25, samples_per_px=8):
    """Boolean (n, n) mask of pixels visited by an Archimedean spiral centre->edge."""
    c = (n - 1) / 2.0
    rmax = 1.025 * n / 2.0                               # slight overshoot, clipped at the frame
    b = rmax / (2 * np.pi * turns)                      # r = b * phi
    phi_max = 2 * np.pi * turns
    # arc length s(phi) = b/2 (phi sqrt(1+phi^2) + asinh(phi)); sample uniformly in s
    s = lambda p: 0.5 * b * (p * np.sqrt(1 + p * p) + np.arcsinh(p))
    total = s(phi_max)
    ns = int(np.ceil(total * samples_per_px))
    phis = np.linspace(0, phi_max, 200001)
    phi = np.interp(np.linspace(0, total, ns), s(phis), phis)
    r = b * phi
    i = np.rint(c + r * np.cos(phi)).astype(int)
    j = np.rint(c + r * np.sin(phi)).astype(int)
    ok = (i >= 0) & (i < n) & (j >= 0) & (j < n)
    m = np.zeros((n, n), dtype=bool)
    m[i[ok], j[ok]] = True
    return m


def spiral_scan(n, noise_sd=0.01, seed=0):
    """C2 / headline / C5 input: sparse (n, n) image, NaN off-trajectory."""
    m = spiral_mask(n)
    f = three_gaussians((n, n))
    rng = np.random.RandomState(seed)
    R = np.full((n, n), np.nan)
    R[m] = f[m] + noise_sd * rng.randn(int(m.sum()))
    return R


def hyperspectral(shape=(64, 64, 16), keep=0.3, seed=0):
    """C3 input: whole-spectrum removal at random xy (corrupt_image3d semantics, gprutils.py:314-359)."""
    e1, e2, e3 = shape
    rng = np.random.RandomState(seed)
    drop = ~(rng.rand(e1 * e2) >= 1.0 - keep)            # RandomState(0).rand(4096) >= 0.7 kept
    x, y = np.meshgrid(np.arange(e1) / e1, np.arange(e2) / e2, indexing="ij")
    z = np.arange(e3)[None, None, :]
    centre = e3 / 2 + 3 * np.sin(2 * np.pi * x)[..., None] * np.cos(2 * np.pi * y)[..., None]
    ampl = 1 + 0.5 * np.cos(2 * np.pi * (x + y))[..., None]
    R = ampl / (1 + ((z - centre) / 2.5) ** 2)
    R = R + 0.01 * rng.randn(*R.shape)
    R = R.reshape(e1 * e2, e3)
    R[drop, :] = np.nan
    return R.reshape(shape)


def dummy_blob(n=32, knockouts=512, seed=0):
    """C1 input: the reference's test fixture shape (test/test_gpreg.py:9-21) at n x n."""
    h = 5
    xx, yy = np.meshgrid(np.arange(0, n * h, h), np.arange(0, n * h, h))
    Z = np.exp(-((xx - 25) ** 2 + (yy - 50) ** 2) / 300.0)
    rng = np.random.RandomState(seed)
    for _ in range(knockouts):
        i = rng.randint(Z.shape[0]); j = rng.randint(Z.shape[1])
        Z[i, j] = np.nan
    return Z


def bo_trial_func(n=128):
    """C4 target: 3-Gaussian trial_func with coordinates scaled by n/25."""
    sc = n / 25.0

    def f(idx):
        x, y = idx[0] / sc, idx[1] / sc

        def g(x0, y0, a, b, fwhm):
            return np.exp(-4 * np.log(2) * (a * (x - x0) ** 2 + b * (y - y0) ** 2) / fwhm ** 2)
        return g(5, 10, 1, 1, 4.5) + g(10, 8, 0.75, 1.5, 7) + g(18, 18, 1, 1.5, 10)
    return f


def make_workload(name, dense=1):
    """BASELINE.json configs by name -> dict(R, kernel, theta(list), d, label, Xfull, jitter).
    `dense` multiplies the number of X_full rows (weak scaling of the dense grid).
    c2 / h512 / c5: 256 / 512 / 1024 spiral scans (RBF); c3: 64x64x16 hyperspectral (Matern52, d = 3);
    c1k: a 128 x 128 spiral (N ~ 3.4 k) for quick checks."""
    ft = FIXED_THETA
    if name in ("c2", "h512", "c5", "c1k"):
        n = {"c2": 256, "h512": 512, "c5": 1024, "c1k": 128}[name]
        R = spiral_scan(n)
        theta = [ft["variance"], ft["noise"], 1.0, ft["lengthscale"], ft["lengthscale"]]
        kern = "RBF"
        label = f"2D {n}x{n} sparse spiral scan, RBF, fixed theta"
    elif name == "c3":
        R = hyperspectral((64, 64, 16))
        theta = [ft["variance"], ft["noise"], 1.0, ft["lengthscale"], ft["lengthscale"], ft["lengthscale_z"]]
        kern = "Matern52"
        label = "3D 64x64x16 hyperspectral, Matern52, fixed theta"
    else:
        raise SystemExit(f"unknown workload {name}")
    sl = [slice(0, R.shape[0], 1.0 / dense)] + [slice(0, e, 1.0) for e in R.shape[1:]]
    Xfull = np.array(np.mgrid[tuple(sl)])                     # gprutils.get_full_grid layout (c, *dims)
    return {"name": name, "R": R, "kernel": kern, "theta": theta, "d": R.ndim, "label": label,
            "Xfull": Xfull, "jitter": ft["jitter"]}


def rows_of(Xgrid):
    """(c, *dims) -> (prod(dims), c): the row layout of gprutils.prepare_test_data (gprutils.py:82)."""
    return Xgrid.reshape(Xgrid.shape[0], -1).T


def sample_rows(M, m):
    """The m grid rows (evenly spread over the M of X_full) the bounded CPU legs and the full-size fixtures use."""
    return np.linspace(0, M - 1, m).astype(np.int64)


if __name__ == "__main__":
    for n in (128, 256, 512, 1024):
        print(n, int(spiral_mask(n).sum()))
    print("C3", int((~np.isnan(hyperspectral())).sum()), "C1", int((~np.isnan(dummy_blob())).sum()))
