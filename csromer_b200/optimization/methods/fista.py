"""``FISTA`` -- the reconstruction loop (reference: optimization/methods/fista.py:17-108).

``run()`` returns ``(cost, Parameter)`` like the reference, mutates ``gx.F[0].reg`` (lambda ends decayed), leaves
``dataset.model_data`` / ``dataset.residual`` set by the final objective evaluation, and returns a deep copy of
``guess_param`` carrying the solution.

Two execution paths with the same semantics:
  * fused (default whenever ``fx`` is a ``Chi2`` on one of this package's operators and ``gx`` holds a single
    ``L1``): the whole loop for ALL lines of sight runs on the GPU inside ``csr_fista`` -- two tensor-core GEMM
    launches per iteration with the residual, soft threshold, momentum and lambda decay in their epilogues;
  * generic: the reference's Python loop over ``fx`` / ``g_prox`` callables, for any other objective objects.

Extensions (not in the reference): ``step`` (gradient step, default 1.0 = reference; SURVEY.md section 0 fact 11 explains
why 1/lambda_max may be wanted), per-pixel ``lambda`` / ``noise`` arrays for batches, ``sparse_out`` (the solution comes
back as compressed rows -- indices and values of its non-zeros -- instead of a dense slab).
"""
import copy
from dataclasses import dataclass

import numpy as np
import torch

from ... import _device as dv
from ...objectivefunction import L1, Chi2, Fi
from ...transformers.dfts import FT
from ..optimizer import Optimizer


def _decayed_lambda(lam0, noise, iterations):
    """lambda after ``iterations`` passes of fista.py:100-106 (sequential subtraction like the reference)."""
    lam0 = np.asarray(lam0, dtype=np.float64)
    noise = np.asarray(noise, dtype=np.float64)
    if lam0.ndim == 0 and noise.ndim == 0:
        lam = float(lam0)
        for _ in range(int(iterations)):
            new = lam - float(noise)
            if new > 0.0:
                lam = new
            else:
                break
        return lam
    with np.errstate(divide="ignore", invalid="ignore"):
        possible = np.where(noise > 0, np.ceil(lam0 / np.where(noise > 0, noise, 1.0)) - 1, iterations)
    decays = np.minimum(np.maximum(possible, 0), iterations)
    return lam0 - decays * noise


@dataclass(init=True, repr=True)
class FISTA(Optimizer):
    fx: Chi2 = None
    gx: Fi = None
    noise: float = None
    step: float = 1.0
    fused: bool = None  # None = automatic
    sparse_out: bool = False  # fused path: return the solution as compressed rows (``SparseRows``) instead of a dense array
    sparse_capacity: int = None  # ... with lists of this many entries allocated up front (no wait for the count)
    async_cost: bool = False  # torch inputs: return the cost as a device tensor instead of waiting for it (streaming runs)

    # ------------------------------------------------------------------------------------------------------
    def _can_fuse(self):
        if self.fused is False:
            return False
        ok = (isinstance(self.fx, Chi2) and isinstance(getattr(self.fx, "dft_obj", None), FT)
              and hasattr(self.gx, "F") and self.gx.F is not None and len(self.gx.F) == 1
              and isinstance(self.gx.F[0], L1) and self.guess_param is not None and self.guess_param.data is not None)
        # the fused loop hard-codes backward = (1/k) E^H (w/s): an operator configured otherwise (NUFFT1D(normalize=False)
        # scales by 1/n, nufft.py:68-69) must go through the generic loop, which calls the operator's own methods
        if ok and getattr(self.fx.dft_obj, "normalize", True) is False:
            ok = False
        if self.fused and not ok:
            raise ValueError("fused=True needs fx=Chi2(dft_obj=<NDFT1D|NUFFT1D>) and gx=OFunction([L1])")
        return ok

    def run(self):
        if self._can_fuse():
            return self._run_fused()
        cost, x = self._generic_loop(self.guess_param.data, self.F_obj.evaluate, self.fx.calculate_gradient_fista,
                                     self.gx, self.maxiter, self.guess_param.n, self.noise, self.verbose)
        param = copy.deepcopy(self.guess_param)
        param.data = x
        return cost, param

    # ------------------------------------------------------------------------------------------------------
    def _iteration_plan(self, lam0):
        """(iters, noise, per-pixel cap or None, early_exit) following fista.py:59-77"""
        noise, maxiter = self.noise, self.maxiter
        cap = None
        if maxiter is None and noise is not None:
            if np.any(np.isnan(np.asarray(noise, dtype=np.float64))):
                raise ValueError("Noise must be a number")
            noise = np.where(np.asarray(noise, dtype=np.float64) != 0.0, noise, 1e-5)
            counts = np.floor(np.asarray(lam0, dtype=np.float64) / noise).astype(np.int64)
            maxiter = int(np.max(counts))
            if counts.ndim:
                cap = counts
            else:
                noise = float(noise)
            if self.verbose:
                print("Iterations set to " + str(maxiter))
        if noise is None:
            noise = 1e-5
        if maxiter is None:
            raise ValueError("FISTA needs maxiter or noise")
        return int(maxiter), noise, cap

    def _run_fused(self):
        chi2, l1 = self.fx, self.gx.F[0]
        op, ds, wav = chi2.dft_obj, chi2.dft_obj.dataset, chi2.wavelet
        plan = op.device_plan()
        dev = plan.device
        lam0 = l1.reg
        iters, noise, cap = self._iteration_plan(lam0)
        x0 = self.guess_param.data
        like_torch = dv.is_torch(x0)

        if np.all(np.asarray(noise) >= np.asarray(lam0)):
            if self.verbose:
                print("Error, noise cannot be greater than lambda")
            param = copy.deepcopy(self.guess_param)
            return 0.0, param

        data, pix = ds.device_data(dev) if hasattr(ds, "device_data") else dv.pack_planar(ds.data, plan.m, plan.ld2m, dev)
        P = data.shape[0]
        w = op._channel_tensor(ds.w, plan)
        s_host = np.asarray(ds.s, dtype=np.float64)
        s = dv.small_to_device(s_host, torch.float32, dev) if s_host.ndim == 1 else op._channel_tensor(s_host, plan)
        k = dv.per_pixel_vector(ds.k, P, dev, "k")
        engine = None
        if wav is not None:
            engine = wav.prepare(2 * plan.n)
            ncoef, ldx = engine.ncoef, engine.ldc
        else:
            ncoef, ldx = 2 * plan.n, plan.ld2n
        if x0.shape[0] != ncoef:
            raise ValueError("guess_param.data has length %d, expected %d real values" % (x0.shape[0], ncoef))
        # README.md:183-190 starts from F_dirty = dft.backward(dataset.data): when the guess is exactly that tensor (untouched,
        # same dataset), its planar device copy is still at hand -- no transpose back, no second dirty spectrum in the loop
        dirty_planar = None
        if engine is None and like_torch and hasattr(op, "dirty_of") and hasattr(self.guess_param, "complex_source"):
            dirty_planar = op.dirty_of(self.guess_param.complex_source())
            if dirty_planar is not None and tuple(dirty_planar.shape) != (P, ldx):
                dirty_planar = None
        if dirty_planar is not None:
            x = torch.empty((P, ldx), dtype=torch.float32, device=dev)
        else:
            x, xpix = dv.pack_real(x0, ncoef, ldx, dev)
        if x.shape[0] != P:
            if x.shape[0] == 1:
                x = x.expand(P, ldx).contiguous()
            else:
                raise ValueError("guess_param.data holds %d lines of sight, dataset holds %d" % (x.shape[0], P))
        lam_t = dv.per_pixel_vector(lam0, P, dev, "lambda")
        noise_t = dv.per_pixel_vector(noise, P, dev, "noise")
        cap_t = None
        if cap is not None:
            cap_t = dv.small_to_device(np.broadcast_to(cap.reshape(-1), (P,)).astype(np.int32), torch.int32, dev)

        if dirty_planar is not None:
            out = plan.fista(data, w, s, k, x, lam_t, noise_t, iters, step=self.step, norm_factor=chi2.norm_factor,
                             iter_cap=cap_t, use_dirty_as_guess=True, want_dirty=False, f_dirty_in=dirty_planar)
        else:
            out = plan.fista(data, w, s, k, x, lam_t, noise_t, iters, step=self.step, norm_factor=chi2.norm_factor,
                             wavelet=engine, iter_cap=cap_t)

        # side effects the reference leaves behind
        model_dev, resid_dev, like = out["model"], out["residual"], ds.data
        ds.defer_model(lambda: dv.unpack_complex_like(model_dev, plan.m, pix, like),
                       lambda: dv.unpack_complex_like(resid_dev, plan.m, pix, like))
        lam_final = _decayed_lambda(lam0, noise, iters if cap is None else np.minimum(cap, iters))
        l1.reg = lam_final
        if self.async_cost and like_torch:   # no host synchronisation: values stay on the device
            cost_t = out["cost"].double()
            chi_v, l1_v = cost_t[:, 0].reshape(pix), cost_t[:, 1].reshape(pix)
            if np.ndim(lam_final):
                lam_final_dev = torch.as_tensor(np.asarray(lam_final, dtype=np.float64).reshape(pix)).to(dev)
            else:
                lam_final_dev = float(lam_final)
        else:
            cost_t = out["cost"].double().cpu().numpy()
            chi_v = cost_t[:, 0].reshape(pix) if pix else float(cost_t[0, 0])
            l1_v = cost_t[:, 1].reshape(pix) if pix else float(cost_t[0, 1])
            lam_final_dev = None
        # F_obj.evaluate(x) of the reference (fista.py:107, ofunction.py:40-46): every term with ITS OWN reg -- the L1 term
        # F_obj holds is usually the very object gx decayed, but it need not be
        cost = None
        terms = getattr(self.F_obj, "F", None)
        if terms:
            cost = 0.0
            for i, term in enumerate(terms):
                if isinstance(term, Chi2) and term.dft_obj is op and term.wavelet is wav:
                    value = chi_v
                elif isinstance(term, L1):
                    value = l1_v
                else:
                    cost = None     # a term the fused loop did not evaluate: leave it to the objects below
                    break
                reg = lam_final_dev if (lam_final_dev is not None and term is l1) else term.reg
                cost = cost + reg * value
                try:
                    self.F_obj.values[i] = value
                except (AttributeError, IndexError, TypeError):
                    pass
        if cost is None:
            cost = chi2.reg * chi_v + (lam_final if lam_final_dev is None else lam_final_dev) * l1_v
        # deep copy of the guess like the reference (fista.py:37-38), minus its (about to be replaced) data array
        held, self.guess_param._data = self.guess_param._data, None
        try:
            param = copy.deepcopy(self.guess_param)
        finally:
            self.guess_param._data = held
        if self.sparse_out:  # > 99 % of an L1 solution is zero: indices and values of the rest (csrc/sparse.cu)
            rows = dv.SparseRows.from_planar(x, ncoef, pix, capacity=self.sparse_capacity if like_torch else None)
            param.data = rows if like_torch else rows.cpu()
        else:
            xr = dv.unpack_real(x, ncoef, pix, like_torch)
            param.data = xr if like_torch else xr.astype(x0.dtype if np.asarray(x0).dtype.kind == "f" else np.float32)
        self.F_dirty = out["f_dirty"]
        self.device_outputs = out  # planar device tensors of the run: f_dirty, model, residual, cost
        self._device_plan, self._pixel_shape = plan, pix
        return cost, param

    # ------------------------------------------------------------------------------------------------------
    def _generic_loop(self, x, F, fx, g_prox, max_iter, n, noise, verbose):
        if x is None and n is not None:
            x = np.zeros(n, dtype=np.complex64)
        t = 1.0
        z = x.copy()
        if max_iter is None and noise is not None:
            if np.isnan(noise):
                raise ValueError("Noise must be a number")
            if noise == 0.0:
                noise = 1e-5
            max_iter = int(np.floor(g_prox.getLambda() / noise))
            if verbose:
                print("Iterations set to " + str(max_iter))
        if noise is None:
            noise = 1e-5
        if noise >= g_prox.getLambda():
            if verbose:
                print("Error, noise cannot be greater than lambda")
            return 0.0, x
        for it in range(max_iter):
            x_prev = x.copy()
            z = z - self.step * fx(z)
            lam = g_prox.getLambda()
            if self.step != 1.0:
                g_prox.setLambda(reg=lam * self.step)
            x = g_prox.calc_prox(z)
            if self.step != 1.0:
                g_prox.setLambda(reg=lam)
            t_prev = t
            t = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t * t))
            z = x + ((t_prev - 1.0) / t) * (x - x_prev)
            if verbose and it % 10 == 0:
                print("Iteration: ", it, " objective function value: {0:0.5f}".format(F(x)))
            lowered = g_prox.getLambda() - noise
            if lowered > 0.0:
                g_prox.setLambda(reg=lowered)
            else:
                if verbose:
                    print("Exit due to negative regularization parameter")
                break
        return F(x), x
