"""pssr2_b200 -- B200-native (sm_100a) implementation of PSSR2's test/predict hot path.

Drop-in surface (same names and argument meaning as the reference package ``pssr``):
    pssr2_b200.predict      predict_images, test_metrics            (pssr/predict.py)
    pssr2_b200.crappifiers  Crappifier, MultiCrappifier, AdditiveGaussian, Poisson, SaltPepper
    pssr2_b200.data         SlidingDataset, ImageDataset            (pssr/data.py)
    pssr2_b200.models       ResUNet, RDResUNet                      (pssr/models)
    pssr2_b200.util         reassemble_sheets, normalize_preds, pixel_metric
All compute runs in ``libpssr_b200.so`` (C ABI: include/pssr_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

__all__ = ["__version__", "models", "crappifiers", "data", "predict", "util", "ops", "dist"]
