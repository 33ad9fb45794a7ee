"""
Model-parameter initialisation at the AnnData level (host, once per fit).

Mirrors the interface of reference initialization/initialize.py (function names, argument
meaning, exceptions) so parity runs start from bit-identical parameters; see the
individual functions for the lines they correspond to.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from .._anndata import AnnData, concat
from ..utils import EPSILON, column_sums_sequential, dict_checker, normalize_WH, shape_checker, type_checker, value_checker
from .methods import (
    _INIT_METHODS,
    init_custom,
    init_flat,
    init_nndsvd,
    init_random,
    init_separableNMF,
)

GIVEN_PARAMETERS_STANDARD_NMF = ["asignatures"]
GIVEN_PARAMETERS_CORRNMF = [
    "asignatures",
    "signature_scalings",
    "sample_scalings",
    "signature_embeddings",
    "sample_embeddings",
    "variance",
]


DEFER_MIN_SIZE = 1 << 20  # exposure matrices at least this large are rescaled / clipped on the device (the host pass over a
# 125k x 20 shard cost more than the whole 20-iteration fit at 8 GPUs)


def initialize_mat(data_mat, n_signatures, method="nndsvd", given_signatures_mat=None, _defer=None, **kwargs):
    """(signatures_mat (k,V), exposures_mat (D,k)): method dispatch, given signatures written over the
    first rows, then normalise columns of W and clip both to EPSILON (reference initialize.py:44-119).

    ``_defer`` (a dict, internal): for large custom exposure matrices (and device-drawn random ones) the second half of that last step,
    ``H <- clip(H * colsum(W))``, is left to the device right after the upload (sal_scale_clip_rows); the
    column sums are returned in ``_defer['exposure_scale']`` and the exposures are handed back untouched."""
    value_checker("method", method, _INIT_METHODS)
    init_device = kwargs.pop("_init_device", None)  # internal: torch.device for the NNDSVD family (models' init_device)
    if method == "custom":
        sigs, expo = init_custom(data_mat, n_signatures, **kwargs)
    elif method == "flat":
        sigs, expo = init_flat(data_mat, n_signatures)
    elif method in ("nndsvd", "nndsvda", "nndsvdar"):
        if init_device is not None:
            from .device_nndsvd import init_nndsvd_device

            sigs, expo = init_nndsvd_device(data_mat, n_signatures, method=method, device=init_device, **kwargs)
        else:
            sigs, expo = init_nndsvd(data_mat, n_signatures, method=method, **kwargs)
    elif method == "random":
        if init_device is not None:
            from .device_nndsvd import init_random_device

            sigs, expo = init_random_device(data_mat, n_signatures, device=init_device, **kwargs)
        else:
            sigs, expo = init_random(data_mat, n_signatures, **kwargs)
    else:
        sigs, expo = init_separableNMF(data_mat, n_signatures, **kwargs)

    if given_signatures_mat is not None:
        type_checker("given_signatures_mat", given_signatures_mat, np.ndarray)
        n_given, n_feat = given_signatures_mat.shape
        if n_feat != data_mat.shape[1]:
            raise ValueError("The given signature matrix has a different number of features than the data.")
        if n_given > n_signatures:
            raise ValueError("The given signature matrix contains too many signatures.")
        sigs[:n_given, :] = given_signatures_mat.copy()

    if _defer is not None and (
        (method == "custom" and expo.size >= DEFER_MIN_SIZE) or (method == "random" and init_device is not None)
    ):
        scale = column_sums_sequential(sigs.T)  # same summation order as normalize_WH
        _defer["exposure_scale"] = scale
        return (sigs / scale[:, None]).clip(EPSILON), expo
    W, H = normalize_WH(sigs.T, expo.T)
    W, H = W.clip(EPSILON), H.clip(EPSILON)
    return W.T, H.T


def check_given_asignatures(given_asignatures, adata, n_signatures) -> None:
    """Type / feature / count compatibility of fixed signatures (reference initialize.py:122-155)."""
    type_checker("given_asignatures", given_asignatures, AnnData)
    if given_asignatures.n_vars != adata.n_vars:
        raise ValueError("The given signatures have a different number of features than the data.")
    if not all(given_asignatures.var_names == adata.var_names):
        raise ValueError("The features of the given signatures and the data are not identical.")
    if given_asignatures.n_obs > n_signatures:
        raise ValueError("The number of given signatures exceeds the number of signatures to initialize.")


def initialize_base(adata, n_signatures, method="nndsvd", given_asignatures=None, _defer=None, **kwargs):
    """Signature AnnData (names 'Sig1..', given signatures first, keeping their names) plus the
    exposure matrix (reference initialize.py:158-218)."""
    given_mat = None
    if given_asignatures is not None:
        check_given_asignatures(given_asignatures, adata, n_signatures)
        given_mat = np.asarray(given_asignatures.X)
    sigs, expo = initialize_mat(np.asarray(adata.X), n_signatures, method, given_mat, _defer=_defer, **kwargs)
    asignatures = AnnData(np.ascontiguousarray(sigs))
    asignatures.var_names = adata.var_names
    asignatures.obs_names = [f"Sig{j + 1}" for j in range(n_signatures)]
    if given_asignatures is not None:
        n_given = given_asignatures.n_obs
        asignatures.obs_names = np.roll(np.asarray(asignatures.obs_names), n_given)
        asignatures = concat([given_asignatures, asignatures[n_given:, :]], join="outer")
    return asignatures, expo


def check_given_parameters_standard_nmf(adata, n_signatures, given_parameters) -> None:
    dict_checker("given_parameters", given_parameters, GIVEN_PARAMETERS_STANDARD_NMF)
    if "asignatures" in given_parameters:
        check_given_asignatures(given_parameters["asignatures"], adata, n_signatures)


def initialize_standard_nmf(adata, n_signatures, method="nndsvd", given_parameters=None, _defer=None, **kwargs):
    """Sets ``adata.obsm['exposures']`` and returns the signature AnnData (reference initialize.py:232-255)."""
    given_parameters = {} if given_parameters is None else given_parameters.copy()
    check_given_parameters_standard_nmf(adata, n_signatures, given_parameters)
    asignatures, expo = initialize_base(adata, n_signatures, method, given_parameters.get("asignatures"), _defer=_defer, **kwargs)
    # (device-drawn exposures of a resident sweep are a device tensor and stay one until the state is built)
    adata.obsm["exposures"] = np.ascontiguousarray(expo) if isinstance(expo, np.ndarray) else expo
    return asignatures


def _check_scalings(given, n_expected, name) -> None:
    type_checker(name, given, np.ndarray)
    shape_checker(name, given, (n_expected,))


def _check_embeddings(given, n_expected, dim_expected, name) -> None:
    type_checker(name, given, np.ndarray)
    shape_checker(name, given, (n_expected, dim_expected))


def check_given_parameters_corrnmf(adata, n_signatures, dim_embeddings, given_parameters: dict[str, Any]) -> None:
    """Reference initialize.py:277-316."""
    dict_checker("given_parameters", given_parameters, GIVEN_PARAMETERS_CORRNMF)
    if "asignatures" in given_parameters:
        check_given_asignatures(given_parameters["asignatures"], adata, n_signatures)
    if "signature_scalings" in given_parameters:
        _check_scalings(given_parameters["signature_scalings"], n_signatures, "given_signature_scalings")
    if "sample_scalings" in given_parameters:
        _check_scalings(given_parameters["sample_scalings"], adata.n_obs, "given_sample_scalings")
    if "signature_embeddings" in given_parameters:
        _check_embeddings(
            given_parameters["signature_embeddings"], n_signatures, dim_embeddings, "given_signature_embeddings"
        )
    if "sample_embeddings" in given_parameters:
        _check_embeddings(given_parameters["sample_embeddings"], adata.n_obs, dim_embeddings, "given_sample_embeddings")
    if "variance" in given_parameters:
        type_checker("given_variance", given_parameters["variance"], [float, int])
        if given_parameters["variance"] <= 0.0:
            raise ValueError("The variance has to be a positive real number.")


def initialize_corrnmf(
    adata,
    n_signatures,
    dim_embeddings,
    method="nndsvd",
    given_parameters=None,
    initialize_sample_embeddings=True,
    **kwargs,
):
    """Signatures as for standard NMF (exposures discarded); scalings start at 0, embeddings ~ N(0, I)
    from the global numpy RNG, variance 1 -- each replaced by its given value (reference initialize.py:319-384).
    """
    if method == "custom":
        raise ValueError(
            "Custom parameter initializations are currently not supported for (multimodal) correlated NMF."
        )
    given_parameters = {} if given_parameters is None else given_parameters.copy()
    check_given_parameters_corrnmf(adata, n_signatures, dim_embeddings, given_parameters)
    asignatures, _ = initialize_base(adata, n_signatures, method, given_parameters.get("asignatures"), **kwargs)

    asignatures.obs["scalings"] = given_parameters.get("signature_scalings", np.zeros(n_signatures))
    adata.obs["scalings"] = given_parameters.get("sample_scalings", np.zeros(adata.n_obs))

    if "signature_embeddings" in given_parameters:
        asignatures.obsm["embeddings"] = given_parameters["signature_embeddings"]
    else:
        asignatures.obsm["embeddings"] = np.random.multivariate_normal(
            np.zeros(dim_embeddings), np.identity(dim_embeddings), size=n_signatures
        )
    if initialize_sample_embeddings:
        if "sample_embeddings" in given_parameters:
            adata.obsm["embeddings"] = given_parameters["sample_embeddings"]
        else:
            adata.obsm["embeddings"] = np.random.multivariate_normal(
                np.zeros(dim_embeddings), np.identity(dim_embeddings), size=adata.n_obs
            )
    variance = float(given_parameters["variance"]) if "variance" in given_parameters else 1.0
    return asignatures, variance


def check_given_parameters_mmcorrnmf(mdata, ns_signatures, dim_embeddings, given_parameters: dict[str, Any]) -> None:
    """Reference initialize.py:387-416: modality names, 'sample_embeddings' and 'variance' are the valid top-level
    keys; the shared parameters cannot be given on the modality level (KeyError)."""
    valid_keys = list(mdata.mod.keys()) + ["sample_embeddings", "variance"]
    dict_checker("given_parameters", given_parameters, valid_keys)
    for (mod_name, adata), n_signatures in zip(mdata.mod.items(), ns_signatures):
        given_mod = given_parameters.get(mod_name, {})
        check_given_parameters_corrnmf(adata, n_signatures, dim_embeddings, given_mod)
        if "sample_embeddings" in given_mod:
            raise KeyError(
                "The sample embeddings are shared across modalities in multimodal correlated NMF. "
                "They cannot be provided as given parameters on the modality level."
            )
        if "variance" in given_mod:
            raise KeyError(
                "The variance parameter of multimodal correlated NMF is shared across modalities. "
                "It cannot be provided as a given parameter on the modality level."
            )


def initialize_mmcorrnmf(mdata, ns_signatures, dim_embeddings, method="nndsvd", given_parameters=None, **kwargs):
    """Per-modality signatures / scalings / signature embeddings as for CorrNMF, new signature names prefixed with the
    modality name, joint sample embeddings ~ N(0, I) in ``mdata.obsm['embeddings']`` (reference initialize.py:419-480)."""
    given_parameters = {} if given_parameters is None else given_parameters.copy()
    check_given_parameters_mmcorrnmf(mdata, ns_signatures, dim_embeddings, given_parameters)
    asignatures = {}
    for (mod_name, adata), n_signatures in zip(mdata.mod.items(), ns_signatures):
        given_mod = given_parameters.get(mod_name, {})
        asigs, _ = initialize_corrnmf(
            adata, n_signatures, dim_embeddings, method, given_mod, initialize_sample_embeddings=False, **kwargs
        )
        n_given = given_mod["asignatures"].n_obs if "asignatures" in given_mod else 0
        names = list(asigs.obs_names)
        asigs.obs_names = names[:n_given] + [f"{mod_name} {name}" for name in names[n_given:]]
        asignatures[mod_name] = asigs
    if "sample_embeddings" in given_parameters:
        mdata.obsm["embeddings"] = given_parameters["sample_embeddings"]
    else:
        mdata.obsm["embeddings"] = np.random.multivariate_normal(
            np.zeros(dim_embeddings), np.identity(dim_embeddings), size=mdata.n_obs
        )
    variance = float(given_parameters["variance"]) if "variance" in given_parameters else 1.0
    return asignatures, variance
