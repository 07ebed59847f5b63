"""get_architect of the 3-stage system (basic_vqa/architect_factory.py:5-16)."""
from argparse import Namespace

import config
from pcdarts.architect_lct import ArchitectLct
from pcdarts.architect_vqa import Architect


def get_architect(ef_model, w_model, ef_optimizer, w_optimizer, reducer=None):
    if config.ARCH_TYPE == 'fixed':
        return None
    if config.ARCH_TYPE == 'darts':
        if config.SKIP_STAGE2:
            # basic_vqa's Architect(model) reads its Adam settings from config (basic_vqa/pcdarts/architect.py:20-22)
            return Architect(ef_model, Namespace(arch_learn_rate=config.ARCH_LEARNING_RATE, arch_wt_decay=config.ARCH_WEIGHT_DECAY,
                                                 qst_only=False), reducer=reducer)
        return ArchitectLct(ef_model, w_model, ef_optimizer, w_optimizer, reducer=reducer)
    raise AssertionError('unrecognized ARCH_TYPE')
