"""
Generate tests/golden/trajectories/*.npz by running the LIVE reference (parklab/Salamander
v0.4.2 mounted at /root/reference) in the build container.  The reference has no test that
pins a multi-iteration trajectory, so these files are the known answers for
the north-star criteria (fp64: objective history within 1e-9 relative; fp32: final KL within
1e-4 relative and signature cosine >= 0.9999).

    python -m oracle.make_golden            # writes the fixtures (a few minutes, CPU)

Each file stores the inputs needed to start from the identical point (W0, H0 after the
reference's own initialisation) plus the reference's history / final parameters.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "trajectories")
DATA = os.path.join(ROOT, "salamander_b200", "data")


def _adata(name="pcawg_breast_sbs.csv", n_samples=None):
    counts = pd.read_csv(os.path.join(DATA, name), index_col=0)
    df = counts.T if n_samples is None else counts.T.iloc[:n_samples]
    return rl.RefAnnData(df)


def klnmf_case(tag, k, seed, fitting_kwargs=None, n_given=0, **ctor):
    sal = rl.load_package()
    adata = _adata()
    model = sal.models.KLNMF(n_signatures=k, init_method="random", **ctor)
    given = None
    if n_given:
        g = _adata()[:n_given, :]
        g.X = g.X / g.X.sum(axis=1, keepdims=True)
        given = {"asignatures": g}
    # replay the reference's own initialisation to record the starting point
    model._setup_adata(adata)
    model._initialize(given, {"seed": seed})
    W0 = np.array(model.asignatures.X)
    H0 = np.array(model.adata.obsm["exposures"])
    adata2 = _adata()
    model.fit(adata2, given_parameters=given, init_kwargs={"seed": seed}, fitting_kwargs=fitting_kwargs)
    fk = fitting_kwargs or {}
    np.savez_compressed(
        os.path.join(OUT, f"{tag}.npz"),
        W0=W0,
        H0=H0,
        history=np.array(model.history["objective_function"]),
        W=np.array(model.asignatures.X),
        H=np.array(model.adata.obsm["exposures"]),
        seed=seed,
        k=k,
        n_given=n_given,
        weights_kl=np.array(fk.get("weights_kl")) if fk.get("weights_kl") is not None else np.zeros(0),
        weights_lhalf=np.array(fk.get("weights_lhalf")) if fk.get("weights_lhalf") is not None else np.zeros(0),
        **{f"ctor_{a}": b for a, b in ctor.items()},
    )
    print(tag, "iterations*freq:", len(model.history["objective_function"]), "final:", model.history["objective_function"][-1])


def mvnmf_case(tag, k, seed, **ctor):
    sal = rl.load_package()
    adata = _adata()
    model = sal.models.MvNMF(n_signatures=k, init_method="random", **ctor)
    model._setup_adata(adata)
    model._initialize(None, {"seed": seed})
    W0 = np.array(model.asignatures.X)
    H0 = np.array(model.adata.obsm["exposures"])
    model.fit(_adata(), init_kwargs={"seed": seed})
    np.savez_compressed(
        os.path.join(OUT, f"{tag}.npz"),
        W0=W0,
        H0=H0,
        history=np.array(model.history["objective_function"]),
        W=np.array(model.asignatures.X),
        H=np.array(model.adata.obsm["exposures"]),
        gamma=model._gamma,
        seed=seed,
        k=k,
        **{f"ctor_{a}": b for a, b in ctor.items()},
    )
    print(tag, "history:", len(model.history["objective_function"]), "final:", model.history["objective_function"][-1], "gamma", model._gamma)


def corrnmf_case(tag, k, dim, seed, n_iter):
    """CorrNMFDet of the live reference: the state right after its own initialisation and ``n_iter`` whole iterations
    (ELBO after each one, final parameters)."""
    sal = rl.load_package()
    kw = dict(n_signatures=k, dim_embeddings=dim, init_method="random", min_iterations=n_iter, max_iterations=n_iter, conv_test_freq=1)
    model = sal.models.CorrNMFDet(**kw)
    adata = _adata()
    model._setup_adata(adata)
    np.random.seed(seed)
    model._initialize(None, {"seed": seed})
    start = dict(
        W0=np.array(model.asignatures.X),
        a0=np.array(model.asignatures.obs["scalings"].values, dtype=float),
        b0=np.array(adata.obs["scalings"].values, dtype=float),
        L0=np.array(model.asignatures.obsm["embeddings"]),
        U0=np.array(adata.obsm["embeddings"]),
        var0=float(model.variance),
    )
    model2 = sal.models.CorrNMFDet(**kw)
    adata2 = _adata()
    np.random.seed(seed)
    model2.fit(adata2, init_kwargs={"seed": seed})
    np.savez_compressed(
        os.path.join(OUT, f"{tag}.npz"),
        **start,
        history=np.array(model2.history["objective_function"]),
        W=np.array(model2.asignatures.X),
        a=np.array(model2.asignatures.obs["scalings"].values, dtype=float),
        b=np.array(adata2.obs["scalings"].values, dtype=float),
        L=np.array(model2.asignatures.obsm["embeddings"]),
        U=np.array(adata2.obsm["embeddings"]),
        H=np.array(adata2.obsm["exposures"]),
        var=float(model2.variance),
        k=k, dim=dim, seed=seed, n_iter=n_iter,
    )
    print(f"{tag}: ELBO history {model2.history['objective_function']}")


def mmcorrnmf_case(tag, ns, dim, seed, n_iter):
    """MultimodalCorrNMF of the live reference on the three PCAWG breast modalities (counts clipped to EPSILON first: the
    multimodal model does not clip and some samples have no SV at all): start state and ``n_iter`` whole iterations."""
    sal = rl.load_package()
    eps = float(np.finfo(np.float32).eps)

    def mdata():
        mods = {}
        for name in ("sbs", "indel", "sv"):
            df = pd.read_csv(os.path.join(DATA, f"pcawg_breast_{name}.csv"), index_col=0).T.astype(float).clip(lower=eps)
            mods[name] = rl.RefAnnData(df)
        return rl.RefMuData(mods)

    kw = dict(ns_signatures=ns, dim_embeddings=dim, init_method="random", min_iterations=n_iter, max_iterations=n_iter, conv_test_freq=1)
    model = sal.models.MultimodalCorrNMF(**kw)
    md0 = mdata()
    model._setup_mdata(md0)
    np.random.seed(seed)
    model._initialize(None, {"seed": seed})
    out = {"U0": np.array(md0.obsm["embeddings"]), "var0": float(model.variance)}
    for name in model.mod_names:
        a_, s_ = md0[name], model.asignatures[name]
        out[f"{name}_W0"], out[f"{name}_L0"] = np.array(s_.X), np.array(s_.obsm["embeddings"])
        out[f"{name}_a0"] = np.array(s_.obs["scalings"].values, dtype=float)
        out[f"{name}_b0"] = np.array(a_.obs["scalings"].values, dtype=float)
    model2 = sal.models.MultimodalCorrNMF(**kw)
    md1 = mdata()
    np.random.seed(seed)
    model2.fit(md1, init_kwargs={"seed": seed})
    out.update(history=np.array(model2.history["objective_function"]), U=np.array(md1.obsm["embeddings"]), var=float(model2.variance))
    for name in model2.mod_names:
        a_, s_ = md1[name], model2.asignatures[name]
        out[f"{name}_W"], out[f"{name}_L"] = np.array(s_.X), np.array(s_.obsm["embeddings"])
        out[f"{name}_a"] = np.array(s_.obs["scalings"].values, dtype=float)
        out[f"{name}_b"] = np.array(a_.obs["scalings"].values, dtype=float)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), ns=np.array(ns), dim=dim, seed=seed, n_iter=n_iter, **out)
    print(f"{tag}: ELBO history {model2.history['objective_function']}")


def init_cases(tag="init_pcawg"):
    """W0 / H0 of the live reference's ``initialize_mat`` for every initialisation method (bit-for-bit targets of
    salamander_b200.initialization)."""
    import importlib

    sal = rl.load_package()
    ref_init = importlib.import_module(sal.__name__ + ".initialization.initialize")
    X = np.asarray(_adata().X, dtype=float).clip(np.finfo(np.float32).eps)
    out = {}
    for method, seed in (("random", 3), ("nndsvd", 2), ("nndsvda", 2), ("nndsvdar", 2), ("flat", None), ("separableNMF", 5)):
        for k in (3, 7):
            kw = {} if seed is None else {"seed": seed}
            W0, H0 = ref_init.initialize_mat(X.copy(), k, method, **kw)
            out[f"{method}_k{k}_W"], out[f"{method}_k{k}_H"] = W0, H0
            out[f"{method}_k{k}_seed"] = -1 if seed is None else seed
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{tag}.npz"), **out)
    print(f"{tag}: {len(out) // 3} initialisations")


def main():
    if not rl.available():
        raise SystemExit("live reference not mounted; nothing to do")
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(7)
    D = 192
    # C1: KLNMF k=5 on PCAWG breast SBS, defaults (runs to convergence, ~7k iterations)
    klnmf_case("klnmf_pcawg_k5_seed0", 5, 0)
    # weighted / l-half / given-signature variants, bounded length
    klnmf_case(
        "klnmf_pcawg_k4_weights",
        4,
        1,
        fitting_kwargs={"weights_kl": rng.uniform(0.5, 2.0, D), "weights_lhalf": rng.uniform(0.0, 4.0, D)},
        min_iterations=300,
        max_iterations=300,
    )
    klnmf_case("klnmf_pcawg_k6_given2", 6, 2, n_given=2, min_iterations=400, max_iterations=400)
    # k % 4 == 0, unweighted: the shapes the tcgen05 (tf32) flavour of the pass covers
    klnmf_case("klnmf_pcawg_k8_seed5", 8, 5)
    klnmf_case("klnmf_pcawg_k4_seed6", 4, 6, min_iterations=1000, max_iterations=1000)
    # C2: MvNMF k=10 on the same data, lam = delta = 1
    mvnmf_case("mvnmf_pcawg_k10_seed0", 10, 0, min_iterations=600, max_iterations=600)
    mvnmf_case("mvnmf_pcawg_k3_lam50", 3, 3, lam=50.0, delta=0.5, min_iterations=300, max_iterations=300)
    # CorrNMFDet (config 4's model on the PCAWG SBS counts): whole iterations incl. both Newton-CG embedding updates
    corrnmf_case("corrnmf_pcawg_k4_dim3_seed3", 4, 3, 3, 6)
    corrnmf_case("corrnmf_pcawg_k6_dim2_seed8", 6, 2, 8, 4)
    mmcorrnmf_case("mmcorrnmf_pcawg_ns322_dim2_seed5", [3, 2, 2], 2, 5, 4)
    init_cases()


if __name__ == "__main__":
    main()
