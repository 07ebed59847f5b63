"""One search step of darts_vqa's Experiment.train (experiment.py:169-200): the alpha-step every
`arch_update_freq` batches, then the w-step — without the dataset / logging / metric glue.

    step = SearchStep(model, architect, optimizer, reducer)
    loss = step.step(train_batch, valid_batch, lr, unrolled=True)
"""
import torch
import torch.nn as nn


class SearchStep:
    def __init__(self, model, architect, optimizer, reducer=None, grad_clip=5.0, qst_only=False):
        self.model = model
        self.architect = architect
        self.optimizer = optimizer
        self.reducer = reducer
        self.grad_clip = grad_clip
        self.qst_only = qst_only
        self.criterion = nn.CrossEntropyLoss()
        self._params = list(model.parameters())

    def w_step(self, image, question, label):
        """experiment.py:187-200."""
        self.optimizer.zero_grad()
        ans_out, qst_out = self.model(image, question)
        qst_loss = self.criterion(qst_out[:, :-1].flatten(end_dim=1), question[:, 1:].flatten())
        loss = qst_loss if self.qst_only else self.criterion(ans_out, label) + qst_loss
        loss.backward()
        if self.reducer is not None:
            self.reducer([p.grad for p in self._params if p.grad is not None])
        nn.utils.clip_grad_norm_(self._params, self.grad_clip)
        self.optimizer.step()
        return loss.detach()

    def alpha_step(self, train_batch, valid_batch, lr, unrolled=True):
        """experiment.py:176-185."""
        self.architect.step(*train_batch, *valid_batch, lr, None, unrolled=unrolled)

    def step(self, train_batch, valid_batch, lr, unrolled=True):
        self.alpha_step(train_batch, valid_batch, lr, unrolled)
        return self.w_step(*train_batch)
