# -*- coding: UTF-8 -*-
"""
B200-native (sm_100a) SF/GPI hot path behind the API of okgarces/deep-successor-features-for-transfer.

    from deep_successor_features_for_transfer_b200.sfdqn import DeepSF, SFDQN, ReplayBuffer      # source/sfdqn.py
    from deep_successor_features_for_transfer_b200.tsfdqn import DeepTSF, TSFDQN                 # source/tsfdqn.py
    from deep_successor_features_for_transfer_b200.ensemble import DeepSF as DeepSFEnsemble      # features/deep.py (G1)
"""
from . import _lib                      # noqa: F401
from .library import PackedSFLibrary, NetSpec   # noqa: F401

__all__ = ['PackedSFLibrary', 'NetSpec']
