"""Shared test helpers: rebuild golden catalogs from their specs (same code path as make_golden.py)."""
import numpy as np

from robot_ebert_b200 import synth


def build_catalog_f32(spec) -> np.ndarray:
    """The fp32 values handed to the catalog store (before dtype quantisation)."""
    if "matrix" in spec:
        return np.asarray(spec["matrix"], dtype=np.float32)
    m = synth.catalog_rows_f32(spec["seed"], 0, spec["n"], spec["d"], spec.get("scale_rows", False))
    for dst, src in spec.get("dup_rows", []):
        m[dst] = m[src]
    for r in spec.get("zero_rows", []):
        m[r] = 0.0
    return m


def build_catalog_f64(spec) -> np.ndarray:
    """The values the catalog of spec['dtype'] stores, upcast to float64 (the oracle's input)."""
    return synth.quantise(build_catalog_f32(spec), spec["dtype"])
