"""
GP-based Bayesian optimisation over image-like grids: drop-in for the reference's
gpim/gpbayes/boptim.py ``boptimizer`` (:22-485).  Host control flow (which point to measure next,
bookkeeping of visited points, checkpoints) stays in Python; the surrogate's training, the dense
predict and -- for the built-in 'cb' / 'ei' / 'poi' -- the acquisition sweep and the top-k
selection run on the device.
"""
import copy
import types

import numpy as np
import torch
from scipy import spatial

from .. import gprutils
from .._lib import ACQ_IDS
from ..gpreg import gpr
from . import acqfunc


class boptimizer:
    """
    Args (as the reference, boptim.py:22-166):
        X_seed (ndarray): sparse grid indices (c, N, M[, L]), NaN where nothing is measured yet
        y_seed (ndarray): sparse observations (N, M[, L]), NaN where nothing is measured yet
        X_full (ndarray): full grid indices
        target_function (callable): takes a tuple of indices, returns the measured value
        acquisition_function: 'cb', 'ei', 'poi' or a function(gpmodel, X_full, X_sparse) -> (acq, (mean, sd))
        exploration_steps, batch_size, batch_update, kernel, lengthscale, sparse, indpoints,
        gp_iterations, seed
        **alpha, **beta, **xi, **use_gpu, **precision, **jitter, **isotropic, **mask, **dscale,
        **batch_dscale, **batch_out_max, **gamma, **memory, **exit_strategy, **extent,
        **save_checkpoints, **filename, **verbose, **learning_rate, **simulate_measurement, **y_true
    """

    def __init__(self, X_seed, y_seed, X_full, target_function, acquisition_function='cb', exploration_steps=10,
                 batch_size=100, batch_update=False, kernel='RBF', lengthscale=None, sparse=False, indpoints=None,
                 gp_iterations=1000, seed=0, **kwargs):
        self.verbose = kwargs.get("verbose", 1)
        self.use_gpu = kwargs.get("use_gpu", False)
        learning_rate = kwargs.get("learning_rate", 5e-2)
        jitter = kwargs.get("jitter", 1.0e-6)
        isotropic = kwargs.get("isotropic", False)
        self.precision = kwargs.get("precision", "double")

        self.surrogate_model = gpr.reconstructor(
            X_seed, y_seed, X_full, kernel, lengthscale, sparse, indpoints, learning_rate, gp_iterations,
            self.use_gpu, self.verbose, seed, isotropic=isotropic, precision=self.precision, jitter=jitter)

        self.X_sparse = X_seed.copy()
        self.y_sparse = y_seed.copy()
        self.X_full = X_full
        self.target_function = target_function
        self.acquisition_function = acquisition_function
        self.exploration_steps = exploration_steps
        self.batch_update = batch_update
        self.batch_size = batch_size
        self.simulate_measurement = kwargs.get("simulate_measurement", False)
        if self.simulate_measurement:
            self.y_true = kwargs.get("y_true")
            if self.y_true is None:
                raise AssertionError("To simulate measurements, add ground truth ('y_true)")
        self.extent = kwargs.get("extent", None)
        self.alpha, self.beta = kwargs.get("alpha", 0), kwargs.get("beta", 1)
        self.xi = kwargs.get("xi", 0.01)
        self.dscale = kwargs.get("dscale", None)
        self.batch_dscale = kwargs.get("batch_dscale", None)
        self.batch_out_max = kwargs.get("batch_out_max", 10)
        self.gamma = kwargs.get("gamma", 0.8)
        self.points_mem = kwargs.get("memory", 10)
        self.exit_strategy = kwargs.get("exit_strategy", 1)
        self.mask = kwargs.get("mask", None)
        self.save_checkpoints = kwargs.get("save_checkpoints", False)
        self.filename = kwargs.get("filename", "./boptim_results")
        self.indices_all, self.vals_all = [], []
        self.target_func_vals, self.gp_predictions = [y_seed.copy()], []

    # ------------------------------------------------------------------------------------------
    def update_posterior(self):
        """Swap the grown training set into the surrogate and retrain (warm theta, fresh Adam);
        boptim.py:239-251."""
        X_new, y_new = gprutils.prepare_training_data(self.X_sparse, self.y_sparse, precision=self.precision)
        self.surrogate_model.model.X = X_new
        self.surrogate_model.model.y = y_new
        self.surrogate_model.train(verbose=self.verbose)

    def evaluate_function(self, indices, y_measured=None):
        """Measure the target at the new point(s) and extend the sparse grid; boptim.py:253-276."""
        indices = [indices] if not self.batch_update else indices
        for idx in indices:
            key = tuple(idx)
            if self.simulate_measurement:
                self.y_sparse[key] = self.y_true[key]
            elif y_measured is not None:
                self.y_sparse[key] = y_measured[key]
            else:
                arg = key if self.extent is None else tuple(i + e[0] for i, e in zip(idx, self.extent))
                self.y_sparse[key] = self.target_function(arg)
        self.X_sparse = gprutils.get_sparse_grid(self.y_sparse, self.extent)
        self.target_func_vals.append(self.y_sparse.copy())

    # ------------------------------------------------------------------------------------------
    def _device_candidates(self):
        """Built-in acquisition: predict + sweep + top-k without leaving the device.
        The incumbent for EI / POI is evaluated at the measured rows only -- the same number the
        reference obtains from nanmax over a full-grid predict of X_sparse (acqfunc.py:57-59,86-88)."""
        sm = self.surrogate_model
        name = self.acquisition_function
        mean, sd = sm.predict(self.X_full, verbose=0)
        mean_d, sd_d = sm._last_pred_device
        mu_best = 0.0
        if name in ("ei", "poi"):
            m_s, s_s = sm.model.predict_sd(sm.model.X)
            mu_best = float(m_s.max().item()) if name == "ei" else float(torch.maximum(m_s.max(), s_s.max()).item())
        mask_d = None
        if self.mask is not None:
            mask_d = torch.as_tensor(np.asarray(self.mask, dtype=np.float64).ravel()).to(mean_d.device, mean_d.dtype)
        k = min(self.batch_size, mean_d.numel(), 1024)
        vals, idx, count, _ = sm.model.engine.acq_sweep(ACQ_IDS[name], mean_d, sd_d, k, mu_best=mu_best, xi=self.xi,
                                                        alpha=self.alpha, beta=self.beta, mask=mask_d)
        # the point filters (checkvalues / update_points) on the ranked list while it is still on the device
        sel = None
        if self.verbose != 2 and mean.ndim <= 4:
            visited = [int(np.ravel_multi_index(tuple(int(c) for c in p), mean.shape)) for p in self.indices_all]
            if self.batch_update:
                bd = (sm.model.kernel.lengthscale.mean().item() if self.batch_dscale is None else self.batch_dscale)
            else:
                bd = 0.0
            sel = sm.model.engine.acq_select(vals, idx, count, mean.shape, visited, memory=self.points_mem,
                                             dscale=0.0 if self.dscale is None else self.dscale, gamma=self.gamma,
                                             batch=self.batch_update, batch_dscale=bd,
                                             batch_out_max=self.batch_out_max if self.batch_update else 0)
        n = int(count.item())
        vals = vals[:n].cpu().numpy().astype(np.float64)
        flat = idx[:n].cpu().numpy()
        indices = np.stack(np.unravel_index(flat, mean.shape), axis=1)
        vals_l, idx_l = vals.tolist(), indices.tolist()
        # only trusted for exactly these lists, and never when a NaN ranks in them (the reference's loops then end
        # through NaN comparisons) or when the list is exhausted (its exit strategies draw from numpy's generator)
        self._sel = None if (sel is None or sel[3] or sel[0] < 0) else {"vals": vals_l, "idx": idx_l, "first": sel[0],
                                                                          "start": sel[1], "picks": sel[2]}
        return vals_l, idx_l, (mean, sd)

    def _host_candidates(self, acq, shape_like):
        self._sel = None
        """Custom acquisition functions (and batch sizes beyond the device top-k): the reference's
        full argsort, boptim.py:303-315."""
        if self.mask is not None:
            acq = self.mask * acq
        order = np.argsort(acq.ravel())
        idx = np.stack(np.unravel_index(order, acq.shape), axis=1)
        vals = acq.ravel()[order]
        if self.mask is not None:
            keep = ~np.isnan(vals)
            vals, idx = vals[keep], idx[:int(keep.sum())]
        return vals[::-1][:self.batch_size].tolist(), idx[::-1][:self.batch_size].tolist()

    def next_point(self):
        """Acquisition over the full grid -> ranked candidate list (values, indices); boptim.py:278-324."""
        if self.verbose:
            print("Computing acquisition function...")
        fn = self.acquisition_function
        if isinstance(fn, str) and fn in ACQ_IDS and self.batch_size <= 1024:
            vals_list, indices_list, pred = self._device_candidates()
        elif isinstance(fn, str) and fn in ACQ_IDS:
            table = {"cb": lambda: acqfunc.confidence_bound(self.surrogate_model, self.X_full, alpha=self.alpha, beta=self.beta),
                     "ei": lambda: acqfunc.expected_improvement(self.surrogate_model, self.X_full, self.X_sparse, xi=self.xi),
                     "poi": lambda: acqfunc.probability_of_improvement(self.surrogate_model, self.X_full, self.X_sparse, xi=self.xi)}
            acq, pred = table[fn]()
            vals_list, indices_list = self._host_candidates(acq, acq)
        elif isinstance(fn, types.FunctionType):
            acq, pred = fn(self.surrogate_model, self.X_full, self.X_sparse)
            vals_list, indices_list = self._host_candidates(acq, acq)
        else:
            raise NotImplementedError("Choose between 'cb', 'ei', and 'poi' acquisition functions or define your own")
        self.gp_predictions.append(pred)
        if not self.batch_update:
            return vals_list, indices_list
        if self.batch_dscale is None:
            batch_dscale_ = self.surrogate_model.model.kernel.lengthscale.mean().item()
        else:
            batch_dscale_ = self.batch_dscale
        return self.update_points(vals_list, indices_list, batch_dscale_)

    def update_points(self, acqfunc_values, indices, dscale):
        """Greedy ball suppression: repeatedly take the best remaining candidate and discard every
        candidate within `dscale` of it (cKDTree), then pad with random candidates up to
        batch_out_max; boptim.py:326-376."""
        sel = getattr(self, "_sel", None)
        if sel is not None and sel["idx"] is indices and sel["vals"] is acqfunc_values and sel["start"] >= 0:
            # gpg_acq_select did the cut and the greedy suppression on the device (same picks, tests/test_gpu_parity.py)
            start = sel["start"]
            vals = np.array(acqfunc_values, dtype=float)[start:]
            pts = np.vstack(indices)[start:]
            chosen_ids = [q - start for q in sel["picks"]]
            chosen_vals = [float(vals[q]) for q in chosen_ids]
            out_idx = pts[chosen_ids].tolist()
            short = self.batch_out_max - len(out_idx)
            if short > 0:
                extra = np.random.randint(0, len(vals), short)
                out_idx.extend(pts[extra].tolist())
                chosen_vals.extend(vals[extra].tolist())
            return chosen_vals, out_idx
        _, val0 = self.checkvalues(indices, acqfunc_values)
        start = int(np.where(np.array(acqfunc_values) == val0)[0][0])
        vals = np.array(acqfunc_values, dtype=float)[start:]
        pts = np.vstack(indices)[start:]
        vals_orig = copy.deepcopy(vals)
        floor = vals.min() - 1
        tree = spatial.cKDTree(pts)
        chosen_vals, chosen_ids = [], []
        best = int(np.argmax(vals))
        while vals[best] > floor:
            chosen_vals.append(vals[best])
            chosen_ids.append(best)
            vals[tree.query_ball_point(pts[best], dscale)] = floor
            best = int(np.argmax(vals))
        chosen_vals = chosen_vals[:self.batch_out_max]
        out_idx = pts[chosen_ids].tolist()[:self.batch_out_max]
        short = self.batch_out_max - len(out_idx)
        if short > 0:
            if self.verbose == 2:
                print("Adding {} random indices".format(short))
            extra = np.random.randint(0, len(vals), short)
            out_idx.extend(pts[extra].tolist())
            chosen_vals.extend(vals_orig[extra].tolist())
        return chosen_vals, out_idx

    def checkvalues(self, idx_list, val_list):
        """First candidate that was not measured before and that keeps the distance
        dscale * gamma**k from the k-th most recent pick (k < memory); boptim.py:378-429."""
        sel = getattr(self, "_sel", None)
        if sel is not None and sel["idx"] is idx_list and sel["vals"] is val_list:
            return idx_list[sel["first"]], val_list[sel["first"]]
        dscale_ = 0 if self.dscale is None else self.dscale

        def too_close(idx):
            recent = self.indices_all[-self.points_mem:][::-1]
            return any(not (np.linalg.norm(np.array(idx) - np.array(p)) > dscale_ * self.gamma ** k)
                       for k, p in enumerate(recent))

        j = 0
        if self.verbose == 2:
            print('Acquisition function max value {} at {}'.format(val_list[j], idx_list[j]))
        if len(self.indices_all) == 0:
            return idx_list[j], val_list[j]
        while (idx_list[j] in self.indices_all) or too_close(idx_list[j]):
            if self.verbose == 2:
                print("Finding the next max point...")
            j += 1
            if j == len(idx_list):
                j = np.random.randint(0, len(idx_list)) if self.exit_strategy else -1
                if self.verbose == 2:
                    print('Index out of list. Exiting with acquisition function value {} at {}'.format(
                        val_list[j], idx_list[j]))
                break
            if self.verbose == 2:
                print('Acquisition function max value {} at {}'.format(val_list[j], idx_list[j]))
        return idx_list[j], val_list[j]

    def single_step(self, *args):
        """One exploration step; boptim.py:431-457."""
        e = args[0]
        if self.verbose:
            print("\nExploration step {} / {}".format(e + 1, self.exploration_steps))
        if e == 0 and not getattr(self, "_skip_initial_training", False):
            self.surrogate_model.train()
        vals, inds = self.next_point()
        if not self.batch_update:
            inds, vals = self.checkvalues(inds, vals)
        self.evaluate_function(inds)
        self.update_posterior()
        if isinstance(vals, float):
            self.indices_all.append(inds)
            self.vals_all.append(vals)
        else:
            self.indices_all.extend(inds)
            self.vals_all.extend(vals)

    def run(self):
        """The exploration loop; boptim.py:459-470.  After resume() the loop continues with the step that
        follows the checkpoint (the initial training of step 0 is not repeated)."""
        for i in range(getattr(self, "_first_step", 0), self.exploration_steps):
            self.single_step(i)
            if self.save_checkpoints:
                self.save_results()
        self.save_results()
        if self.verbose:
            print("\nExploration completed")

    def save_results(self, *args):
        """np.save of {gp_pred, func_val, inds_all, vals_all}; boptim.py:472-485."""
        filename = args[0] if args else self.filename
        results = {'gp_pred': self.gp_predictions, 'func_val': self.target_func_vals,
                   'inds_all': np.array(self.indices_all), 'vals_all': np.array(self.vals_all)}
        # beyond the reference's four keys: the surrogate's unconstrained hyper-parameters, so that a run can
        # be resumed exactly where it stopped (the reference writes checkpoints but has no way to load them)
        results['engine_state'] = {'u': self.surrogate_model.model._u.detach().cpu().numpy().copy(),
                                   'steps_done': len(self.gp_predictions),
                                   'trained': len(getattr(self.surrogate_model, "hyperparams", {}).get("noise", ())) > 0,
                                   'np_random_state': np.random.get_state()}
        if hasattr(self.surrogate_model.model, "Xu"):      # sparse surrogate: the trained inducing inputs belong to the state
            results['engine_state']['Xu'] = self.surrogate_model.model.Xu.detach().cpu().numpy().copy()
        np.save(filename + ".npy", results)

    def resume(self, filename=None):
        """Load a checkpoint written by save_results() and continue from it: measured values, pick history,
        stored predictions and the trained hyper-parameters are restored, and run() goes on with the next
        exploration step.  (the reference has no load / resume path.)"""
        filename = self.filename if filename is None else filename
        res = np.load(filename + ".npy", allow_pickle=True).item()
        self.gp_predictions = list(res['gp_pred'])
        self.target_func_vals = list(res['func_val'])
        self.indices_all = np.asarray(res['inds_all']).tolist()
        self.vals_all = np.asarray(res['vals_all']).tolist()
        self.y_sparse = self.target_func_vals[-1].copy()
        self.X_sparse = gprutils.get_sparse_grid(self.y_sparse, self.extent)
        model = self.surrogate_model.model
        X_new, y_new = gprutils.prepare_training_data(self.X_sparse, self.y_sparse, precision=self.precision)
        model.X, model.y = X_new, y_new
        state = res.get('engine_state')
        if state is None:                                  # a reference-format file: retrain from the prior draw
            self.surrogate_model.train(verbose=self.verbose)
        else:
            model.load_unconstrained(state['u'])
            if state.get('Xu') is not None and hasattr(model, "Xu"):
                model.Xu = torch.as_tensor(state['Xu'])
        # continue with the step after the last completed one; a checkpoint written before any step (or a
        # reference-format file without predictions) starts at step 0, INCLUDING its initial training -- unless the
        # checkpoint carries trained hyper-parameters of completed steps
        done = int(state['steps_done']) if state is not None else len(self.gp_predictions)
        self._first_step = done
        self._skip_initial_training = done == 0 and state is not None and bool(state.get('trained', False))
        if state is not None and state.get('np_random_state') is not None:
            np.random.set_state(state['np_random_state'])   # the random fall-backs of checkvalues / update_points
        return self
