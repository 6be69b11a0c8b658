"""Mirror of the reference's `config/config_fingerprint.py` (config_fingerprint.py:1-45, config_fingerprint.yml:1-50).

Same module-level names (`cfg`, `get_path`, `*_DIR`, `DB_CONFIG`, `PREPROCESSING_PARAMS`, `BINARIZATION_PARAMS`,
`ORIENTATION_PARAMS`, `GENERAL_PARAMS`, `print_config_summary`).  The YAML is looked up at `$FPB200_CONFIG_YAML`, then
`./config/config_fingerprint.yml` (where the reference keeps it); without a file the reference's shipped values are used.

In the reference the NUMERIC sections are dead: no hot-path function reads them, every stage uses hard-coded values that
differ from the YAML, so the defaults of this package are those hard-coded values and parity is
judged on them.  The YAML becomes live only on request - `FPB200_YAML_OVERRIDES=1` or an explicit call of
`overrides()` - and then only for the keys a kernel parameter exists for:

    orientation.quality_window / quality_threshold / coherence_threshold / min_distance / margin
        -> `params` of postprocess_minutiae (post_processing.py:77-81; fpb_set_post_params)
    general.rel_threshold -> `rel_thresh` of thinning_and_cleaning (fingerprint_preprocess.py:161; fpb_set_rel_threshold)

Every other numeric key (CLAHE clip, bilateral, Sauvola window / k, patch size, object / hole sizes, block size, sigmas) has
no counterpart in the reference's code either and is reported by `unused_keys()` instead of being silently ignored."""
from __future__ import annotations

import os
from typing import Dict, Optional

BASE_DIR = os.path.abspath(os.getcwd())
CONFIG_YAML = os.environ.get("FPB200_CONFIG_YAML") or os.path.join(BASE_DIR, "config", "config_fingerprint.yml")

# the values of the reference's shipped YAML (used when no file is found)
_SHIPPED = {
    "paths": {"base_dir": "", "dataset_dir": "./dataset", "processed_dir": "./dataset/processed",
              "features_dir": "./data/features", "metadata_dir": "./data/metadata", "debug_dir": "./data/features/debug"},
    "database": {"host": "localhost", "dbname": "biometria", "user": "postgres", "password": "postgres", "port": 5432},
    "preprocessing": {"clahe_clip_limit": 2.0, "clahe_tile_size": 8,
                      "bilateral": {"d": 5, "sigma_color": 50.0, "sigma_space": 7.0}, "gaussian_sigma": 0.7},
    "binarization": {"sauv_wo_win": 25, "sauv_k": 0.2, "local_patch": 64, "min_obj_size": 30, "max_hole_size": 100,
                     "min_segment_area": 5000},
    "orientation": {"quality_window": 25, "quality_threshold": 0.25, "coherence_threshold": 0.3, "min_distance": 10.0,
                    "orientation_window": 15, "margin": 40, "orient_sigma": 10.0},
    "general": {"block_size": 16, "energy_threshold": 0.01, "rel_threshold": 0.2, "vis_scale": 8},
}

# what the hot path really uses (hard-coded in the reference) - the defaults of this package
HARD_CODED = {"post_params": {"quality_window": 25, "quality_threshold": 0.15, "coherence_threshold": 0.2, "min_distance": 8.0,
                              "margin": 30, "max_minutiae": 60, "patch_radius": 15},
              "rel_threshold": 0.1}

_LIVE_ORIENTATION_KEYS = ("quality_window", "quality_threshold", "coherence_threshold", "min_distance", "margin")


def load(path: Optional[str] = None) -> Dict:
    path = path or CONFIG_YAML
    if path and os.path.isfile(path):
        import yaml
        with open(path, "r") as f:
            return yaml.safe_load(f) or {}
    return {k: (dict(v) if isinstance(v, dict) else v) for k, v in _SHIPPED.items()}


cfg = load()


def get_path(key: str, default: str) -> str:
    path = cfg.get("paths", {}).get(key, default)
    return os.path.abspath(os.path.join(BASE_DIR, path))


METADATA_DIR = get_path("metadata_dir", "data/metadata")
DATASET_DIR = get_path("dataset_dir", "./dataset")
SORTED_DATASET_DIR = get_path("sorted_dataset_dir", "./dataset/sorted_dataset")
PROCESSED_DIR = get_path("processed_dir", "./dataset/processed")
FEATURES_DIR = get_path("features_dir", "./data/features")
DEBUG_DIR = get_path("debug_dir", "./data/features/debug")

DB_CONFIG = cfg.get("database", {})
PREPROCESSING_PARAMS = cfg.get("preprocessing", {})
BINARIZATION_PARAMS = cfg.get("binarization", {})
ORIENTATION_PARAMS = cfg.get("orientation", {})
GENERAL_PARAMS = cfg.get("general", {})


def overrides(config: Optional[Dict] = None) -> Dict:
    """{"post_params": {...}, "rel_threshold": x} taken from the YAML sections (only keys present in the file)."""
    c = cfg if config is None else config
    post = {k: c["orientation"][k] for k in _LIVE_ORIENTATION_KEYS if k in (c.get("orientation") or {})}
    if "quality_window" in post:
        qw = int(post["quality_window"])
        if qw < 1 or qw > 33:
            raise ValueError("orientation.quality_window must be in [1, 33] (the density kernel's tile)")
        post["quality_window"] = qw
    out: Dict = {}
    if post:
        out["post_params"] = post
    if "rel_threshold" in (c.get("general") or {}):
        out["rel_threshold"] = float(c["general"]["rel_threshold"])
    return out


def active_overrides() -> Dict:
    """What the drivers apply: {} unless FPB200_YAML_OVERRIDES=1 (defaults must stay the reference's hard-coded values)."""
    return overrides() if os.environ.get("FPB200_YAML_OVERRIDES", "0") == "1" else {}


def unused_keys(config: Optional[Dict] = None):
    """Numeric YAML keys nothing reads - neither here nor in the reference (dotted names)."""
    c = cfg if config is None else config
    live = {f"orientation.{k}" for k in _LIVE_ORIENTATION_KEYS} | {"general.rel_threshold"}
    out = []
    for sec in ("preprocessing", "binarization", "orientation", "general"):
        def walk(prefix, d):
            for k, v in (d or {}).items():
                name = f"{prefix}.{k}"
                if isinstance(v, dict):
                    walk(name, v)
                elif name not in live:
                    out.append(name)
        walk(sec, c.get(sec))
    return out


def print_config_summary():
    print("\n=== CONFIGURAZIONE CARICATA ===")
    print("Percorsi:")
    for k, v in cfg.get("paths", {}).items():
        print(f"  {k}: {get_path(k, v)}")
    print("\nDatabase:")
    for k, v in DB_CONFIG.items():
        print(f"  {k}: {v}")
    print("\nParametri preprocessing:", PREPROCESSING_PARAMS)
    print("Parametri binarizzazione:", BINARIZATION_PARAMS)
    print("Parametri orientazione:", ORIENTATION_PARAMS)
    print("Parametri generali:", GENERAL_PARAMS)
    print("Override attivi (FPB200_YAML_OVERRIDES):", active_overrides() or "nessuno - valori hard-coded del riferimento")
    print("================================\n")
