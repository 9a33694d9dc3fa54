"""Fit-and-predict study over 80/20 splits of the materials datasets.

Mirrors the callable surface of the reference's ``benchmark_predictions.py`` (``fit_predict`` :19-75,
``evaluate_model_across_datasets`` :78-166, the ``__main__`` sweep :169-206 without its plots): every model is fitted
on 80 % of a dataset (n up to 480 on CrossedBarrel) and scored by the mean absolute error of ``posterior.mean`` on the
held-out 20 %.  The reference fits one (model, dataset, seed) at a time; here all seeds of a (model, dataset) pair are
ONE batched launch per step (fit, predict) through the C ABI -- the same kernels as the BO path, with the exact-GP
working square in the global workspace above ``_lib.MAX_N_FIT`` training points (``stp_exact_*_large``) and the
variational global-square path above 144.  There is no CPU fallback.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, priors, varops
from .models import ExactGP, VarGP, VarLGP, VarTGP
from .optim import train_exact_model_botorch, train_natural_variational_model
from .utils import TorchStandardScaler, preprocess_data, set_seeds

_KIND = {ExactGP: "ExactGP", VarGP: "VarGP", VarTGP: "VarTGP", VarLGP: "VarLGP"}


def fit_predict(X_train, y_train, X_test, y_test, model_class, epochs=100, lr=0.01, **kwargs):
    """One model on one split through the model API (benchmark_predictions.py:19-75): returns
    (predictions, ground_truth, mae).  For the Monte-Carlo models ``posterior.mean`` is the [S, Nt] sample tensor and
    the MAE averages over it, exactly as the reference's broadcasting does (:62-73)."""
    X_train = X_train.double()
    y_train = y_train.double().flatten()
    X_test = X_test.double()
    y_test = y_test.double().flatten()
    y_scaler = TorchStandardScaler()
    y_train_scaled = y_scaler.fit_transform(y_train)
    if model_class == ExactGP:
        kwargs.update({"X_train": X_train, "y_train": y_train_scaled})
        model = model_class(**kwargs).double()
        train_exact_model_botorch(model, X_train, y_train_scaled)
    else:
        if "inducing_points" not in kwargs:
            kwargs["inducing_points"] = X_train
        model = model_class(**kwargs).double()
        train_natural_variational_model(model, X_train, y_train_scaled, epochs=epochs, lr=lr, y_standardize=False)
    model.eval()
    with torch.no_grad():
        predictions_scaled = model.posterior(X_test).mean
    predictions = y_scaler.inverse_transform(predictions_scaled.cpu())
    mae = torch.mean(torch.abs(predictions - y_test))
    return predictions, y_test, mae.item()


def split_indices(n_points: int, seed: int, test_size: float = 0.2):
    """(train_indices, test_indices) of benchmark_predictions.py:130-137: ``set_seeds(seed)`` then one
    ``torch.randperm`` from the global generator; the first int(test_size * N) entries are the test set."""
    set_seeds(int(seed))
    perm = torch.randperm(n_points)
    k = int(test_size * n_points)
    return perm[k:], perm[:k]


def _ls_prior(lengthscale_prior):
    if lengthscale_prior is None:
        return priors.VAR_LENGTHSCALE_PRIOR
    if isinstance(lengthscale_prior, (tuple, list)):
        return tuple(float(v) for v in lengthscale_prior)
    return (float(lengthscale_prior.loc), float(lengthscale_prior.scale))      # a gpytorch-style LogNormalPrior


def launch_fit_predict(kind: str, X_train, y_train, X_test, epochs: int = 100, lr: float = 0.01, lengthscale_prior=None,
                       device=None, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Enqueue fit + predict of all splits of one (model kind, dataset) pair on the CURRENT CUDA stream and return the
    DEVICE tensors of fit_predict_batch's result without synchronising (``_keep`` holds the buffers the kernels still
    read).  Callers overlap several pairs by launching each on its own stream (evaluate_models_across_datasets)."""
    if not torch.cuda.is_available():
        raise RuntimeError("fit_predict_batch needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = device or torch.device("cuda", torch.cuda.current_device())
    S_, n, d = X_train.shape
    Xd = X_train.to(dev, torch.double).contiguous()
    yraw = y_train.to(dev, torch.double).contiguous()
    Xq = X_test.to(dev, torch.double).contiguous()
    mu = yraw.mean(1, keepdim=True)
    sd = yraw.std(1, unbiased=False, keepdim=True) + 1e-10
    nn = torch.full((S_,), n, dtype=torch.int32, device=dev)
    pid = torch.arange(S_, dtype=torch.int32, device=dev)
    if kind == "ExactGP":
        ys = ((yraw - mu) / sd).contiguous()
        theta = torch.tensor([priors.exact_theta0(d)] * S_, dtype=torch.double, device=dev)
        lo, hi = priors.exact_bounds(d)
        _, nit, nfev, status = _lib.exact_fit(Xd, ys, theta, lo, hi, priors.exact_prior_vector(d), n=nn)
        out = _lib.exact_acq(Xq, Xd, ys, theta, pool_id=pid, n=nn, want_pointwise=True)
        status = torch.where(status != 0, status, out["status"])
        return {"mean": out["mean"] * sd + mu, "var": out["var"], "status": status, "theta": theta, "nfev": nfev,
                "_keep": (Xd, ys, Xq, nn, pid, out)}
    if kind not in _lib.LIK:
        raise ValueError(f"unknown model kind {kind!r}")
    cap = _lib.var_capacity(n + 1, d)
    ls_prior = _ls_prior(lengthscale_prior)
    Xc = torch.zeros(S_, cap, d, dtype=torch.double, device=dev); Xc[:, :n] = Xd
    yc = torch.zeros(S_, cap, dtype=torch.double, device=dev); yc[:, :n] = yraw
    st = varops.alloc_state(S_, cap, d, dev)
    status = torch.zeros(S_, dtype=torch.int32, device=dev)
    work = _lib.var_workspace(S_, cap, dev)
    _lib.var_fit(Xc, yc, nn, st, _lib.LIK[kind], int(epochs), float(lr), varops.ADAM_LR, varops.var_hyp0(d, ls_prior),
                 varops.var_prior_vector(ls_prior), varops.gauss_hermite_table(), status, work, fresh=True, seed=int(seed))
    bf = torch.zeros(S_, dtype=torch.double, device=dev)
    out = _lib.var_acq(Xq, st, nn, 0, status, work, best_f=bf, pool_id=pid, S=1, want_pointwise=True)
    return {"mean": out["mean"] * sd + mu, "var": out["var"], "status": status, "scale": sd.reshape(-1),
            "_keep": (Xc, yc, Xq, nn, pid, st, work, bf, out)}


def _to_host(res: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: v.cpu() for k, v in res.items() if k != "_keep"}


def fit_predict_batch(kind: str, X_train: torch.Tensor, y_train: torch.Tensor, X_test: torch.Tensor,
                      epochs: int = 100, lr: float = 0.01, lengthscale_prior=None, num_samples: int = 2048,
                      device=None, seed: int = 0) -> Dict[str, torch.Tensor]:
    """All splits of one (model kind, dataset) pair at once.  X_train [S, n, d], y_train [S, n] (raw), X_test [S, m, d].

    Returns host tensors: ``mean`` [S, m] = posterior mean in the ORIGINAL scale of y, ``var`` [S, m] = the latent
    predictive variance in the standardised scale (exact: with the likelihood noise, as ExactGP.posterior), ``status``
    [S] (0 = ok).  The targets are standardised per split (TorchStandardScaler: population std + 1e-10,
    benchmark_predictions.py:46-48); the exact GP is fitted on the standardised targets, the variational fit
    standardises inside the kernel.  ``kind``: "ExactGP" | "VarGP" | "VarTGP" | "VarLGP"."""
    return _to_host(launch_fit_predict(kind, X_train, y_train, X_test, epochs, lr, lengthscale_prior, device, seed))


def _mae(kind: str, res: Dict[str, torch.Tensor], y_test: torch.Tensor, num_samples: int, generator=None) -> torch.Tensor:
    """MAE per split.  ExactGP / VarGP: |mean - y| averaged over the test points.  VarTGP / VarLGP: the reference
    averages |f_s - y| over the S samples f_s = m + sqrt(v) eps of ``posterior.mean`` as well (models.py:187-202,
    benchmark_predictions.py:62-73); the draws here come from ``generator`` (the reference's global-RNG stream is not
    reproducible across implementations)."""
    mean = res["mean"]
    if kind in ("ExactGP", "VarGP"):
        return (mean - y_test).abs().mean(1)
    sdev = res["var"].clamp_min(0.0).sqrt() * res["scale"].reshape(-1, 1)
    eps = torch.randn((num_samples,) + tuple(mean.shape), dtype=torch.double, generator=generator)
    return (mean.unsqueeze(0) + sdev.unsqueeze(0) * eps - y_test.unsqueeze(0)).abs().mean((0, 2))


def evaluate_models_across_datasets(model_classes, n_seeds=10, test_size=0.2, epochs=100, lr=0.01, data_dir: str = "data",
                                    seeds: Optional[Sequence[int]] = None, seeds_file: str = "random_seeds.txt",
                                    dataset_files: Optional[Iterable[str]] = None, verbose: bool = True,
                                    concurrent: bool = True, **model_kwargs) -> Dict[str, Dict[str, dict]]:
    """The study for several model classes at once: {model name: result dictionary of evaluate_model_across_datasets}.
    Every (model, dataset) pair is one batched fit launch + one batched predict launch on its OWN CUDA stream, the
    longest first, so that the pairs overlap on the device (a pair occupies one SM per seed: 25 of 148): the study
    takes as long as its longest pair instead of the sum.  ``concurrent=False`` runs the pairs one after another."""
    kinds = []
    for model_class in model_classes:
        kind = _KIND.get(model_class)
        if kind is None:
            raise NotImplementedError(f"{model_class!r}: only the stock model classes have a batched fit-and-predict path")
        kinds.append(kind)
    unknown = set(model_kwargs) - {"lengthscale_prior"}
    if unknown:
        raise NotImplementedError(f"unsupported model_kwargs {sorted(unknown)}")
    if not torch.cuda.is_available():
        raise RuntimeError("the fit-and-predict study needs a CUDA device (sm_100a); there is no CPU fallback")
    if seeds is None:
        seeds = np.loadtxt(seeds_file, dtype=int)[:n_seeds]
    seeds = [int(s) for s in seeds][:n_seeds]
    files = sorted(dataset_files) if dataset_files is not None else \
        [f for f in os.listdir(data_dir) if f.endswith(".csv")]
    data = {}
    for dataset_file in files:
        name = os.path.splitext(os.path.basename(dataset_file))[0]
        X, y = preprocess_data(filepath=os.path.join(data_dir, dataset_file))
        X, y = X.double(), y.double().flatten()
        if verbose:
            print(f"\nEvaluating dataset: {name}\n{tuple(X.shape)}")
        tr, te = zip(*[split_indices(len(X), s, test_size) for s in seeds])
        data[name] = (X, y, torch.stack(tr), torch.stack(te))
    # longest pairs first: cost ~ n^3 x (epochs for the variational models, ~100 closures for the exact one)
    pairs = sorted(((kind, name) for kind in kinds for name in data),
                   key=lambda p: -(data[p[1]][2].shape[1] ** 3) * (epochs if p[0] != "ExactGP" else 100.0))
    main_stream = torch.cuda.current_stream()
    launched = {}
    for kind, name in pairs:
        X, y, tr, te = data[name]
        st = torch.cuda.Stream() if concurrent else main_stream
        st.wait_stream(main_stream)
        with torch.cuda.stream(st):
            launched[(kind, name)] = (st, launch_fit_predict(kind, X[tr], y[tr], X[te], epochs=epochs, lr=lr,
                                                             lengthscale_prior=model_kwargs.get("lengthscale_prior")))
    out: Dict[str, Dict[str, dict]] = {}
    for model_class, kind in zip(model_classes, kinds):
        results: Dict[str, dict] = {}
        for name in data:
            st, dres = launched.pop((kind, name))
            st.synchronize()
            res = _to_host(dres)
            del dres
            X, y, tr, te = data[name]
            mae = _mae(kind, res, y[te], 2048, torch.Generator().manual_seed(seeds[0]))
            bad = res["status"] != 0
            mae[bad] = float("nan")          # the reference would have raised inside fit_predict for such a split
            results[name] = {
                "seeds": list(seeds), "mae_scores": mae.tolist(),
                "predictions": [res["mean"][i] for i in range(len(seeds))],
                "ground_truth": [y[te[i]] for i in range(len(seeds))],
                "mean_mae": float(np.nanmean(mae.numpy())), "std_mae": float(np.nanstd(mae.numpy())),
                "status": res["status"].tolist(),
            }
        out[model_class.__name__] = results
    return out


def evaluate_model_across_datasets(model_class, n_seeds=10, test_size=0.2, epochs=100, lr=0.01, data_dir: str = "data",
                                   seeds: Optional[Sequence[int]] = None, seeds_file: str = "random_seeds.txt",
                                   dataset_files: Optional[Iterable[str]] = None, verbose: bool = True,
                                   **model_kwargs) -> Dict[str, dict]:
    """benchmark_predictions.py:78-166 with the seeds of a dataset batched: for every ``*.csv`` under ``data_dir`` the
    splits of the first ``n_seeds`` seeds are fitted and scored together (the datasets overlap on the device, see
    evaluate_models_across_datasets).  Returns the reference's result dictionary
    {dataset: {seeds, mae_scores, predictions, ground_truth, mean_mae, std_mae}}."""
    return evaluate_models_across_datasets([model_class], n_seeds, test_size, epochs, lr, data_dir, seeds, seeds_file,
                                           dataset_files, verbose, **model_kwargs)[model_class.__name__]


def main(n_seeds: int = 25, epochs: int = 600, lr: float = 0.01, data_dir: str = "data",
         out_path: str = "results/dimscaled_result_db.npy", model_classes=(VarTGP, VarGP, ExactGP)) -> np.ndarray:
    """The sweep of benchmark_predictions.py:169-195: result_db[model, dataset, seed] = MAE, saved as .npy."""
    datasets: List[str] = [f for f in os.listdir(data_dir) if f.endswith(".csv")]
    result_db = np.zeros((len(model_classes), len(datasets), n_seeds))
    everything = evaluate_models_across_datasets(list(model_classes), n_seeds=n_seeds, test_size=0.2, epochs=epochs, lr=lr,
                                                 data_dir=data_dir, dataset_files=datasets)
    for i, model_class in enumerate(model_classes):
        results = everything[model_class.__name__]
        for j, dataset_name in enumerate(datasets):
            result_db[i, j, :] = results[os.path.splitext(dataset_name)[0]]["mae_scores"]
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    np.save(out_path, result_db)
    return result_db


if __name__ == "__main__":
    main()
