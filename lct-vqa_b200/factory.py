"""Builders with the reference's names (darts_vqa/factory.py:6-35, basic_vqa/architect_factory.py:5-16)."""
import torch.optim as optim
from torch.optim import lr_scheduler

from pcdarts.architect_vqa import Architect
from vqa_model import VqaModel


def get_vqa_model(args, dataset):
    if getattr(args, "unified", False):
        raise NotImplementedError("the unified question+answer decoder is outside the PC-DARTS search path")
    return VqaModel(args.embed_size, dataset.qst_vocab.vocab_size, dataset.ans_vocab.vocab_size,
                    args.word_embed_size, args.num_layers, args.hidden_size, args.arch_type)


def get_optimizer(args, model):
    return optim.Adam(model.parameters(), lr=args.learn_rate)


def get_scheduler(args, optimizer):
    return lr_scheduler.StepLR(optimizer, step_size=args.step_size, gamma=args.gamma)


def get_architect(args, model, reducer=None):
    if args.arch_type == 'vgg':
        return None
    if args.arch_type == 'darts':
        return Architect(model, args, reducer=reducer)
    raise Exception(f'Unrecognized arch_type: {args.arch_type}')
