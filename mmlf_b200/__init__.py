"""mmlf_b200: B200-native (sm_100a) hot path of titus-leistner/mmlf behind the reference's Python API.

Package layout mirrors the reference (``mmlf.model.feed_forward`` -> ``mmlf_b200.model.feed_forward`` ...); the
top-level ``mmlf`` package in this repository re-exports these modules so ``python -m mmlf.train.cli`` works unchanged.
"""
__version__ = '0.1.0'
