"""W model of the 3-stage LCT system (basic_vqa/models.py:7-125): VGG19 image encoder (frozen feature extractor),
question LSTM encoder, element-wise fusion, 1000-way answer head.  Stock PyTorch — it contains no PC-DARTS search
network; it is here because `ArchitectLct` differentiates through it.  Same class / attribute /
state_dict names and registration order as the reference.

`pretrained=None` follows config.PRETRAIN_ENC (the reference hard-codes pretrained=True, models.py:23, which needs a
network download); pass False to build the architecture with random weights.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import config


def softXEnt(pred, target):
    """Soft cross entropy (models.py:7-10)."""
    logprobs = F.log_softmax(pred, dim=1)
    return -(target * logprobs).sum() / pred.shape[0]


def _copy_dropout(src, dst):
    """model.new() builds fresh modules; carry the dropout probabilities over (they are configuration, not state)."""
    probs = {n: m.p for n, m in src.named_modules() if isinstance(m, nn.Dropout)}
    for n, m in dst.named_modules():
        if isinstance(m, nn.Dropout) and n in probs:
            m.p = probs[n]


class ImgEncoder(nn.Module):
    def __init__(self, embed_size, pretrained=None):
        super().__init__()
        import torchvision.models as tvm
        pretrained = config.PRETRAIN_ENC if pretrained is None else pretrained
        model = tvm.vgg19(weights=tvm.VGG19_Weights.IMAGENET1K_V1 if pretrained else None)
        in_features = model.classifier[-1].in_features
        model.classifier = nn.Sequential(*list(model.classifier.children())[:-1])     # drop the ImageNet head
        self.model = model
        self.fc = nn.Linear(in_features, embed_size)

    def forward(self, image):
        with torch.no_grad():                                   # models.py:35-36: fixed feature extractor
            img_feature = self.model(image)
        img_feature = self.fc(img_feature)
        return img_feature.div(img_feature.norm(p=2, dim=1, keepdim=True).detach())


class QstEncoder(nn.Module):
    def __init__(self, qst_vocab_size, word_embed_size, embed_size, num_layers, hidden_size):
        super().__init__()
        self.word2vec = nn.Embedding(qst_vocab_size, word_embed_size)
        self.tanh = nn.Tanh()
        self.lstm = nn.LSTM(word_embed_size, hidden_size, num_layers)
        self.fc = nn.Linear(2 * num_layers * hidden_size, embed_size)

    def forward(self, question):
        qst_vec = self.tanh(self.word2vec(question)).transpose(0, 1)
        _, (hidden, cell) = self.lstm(qst_vec)
        qst_feature = torch.cat((hidden, cell), 2).transpose(0, 1)
        qst_feature = self.tanh(qst_feature.reshape(qst_feature.size(0), -1))
        return self.fc(qst_feature)


class VqaModel(nn.Module):
    def __init__(self, embed_size, qst_vocab_size, ans_vocab_size, word_embed_size, num_layers, hidden_size,
                 pretrained=None):
        super().__init__()
        self.pretrained = pretrained
        self.img_encoder = ImgEncoder(embed_size, pretrained)
        self.qst_encoder = QstEncoder(qst_vocab_size, word_embed_size, embed_size, num_layers, hidden_size)
        self.tanh = nn.Tanh()
        self.dropout = nn.Dropout(0.5)
        self.fc1 = nn.Linear(embed_size, ans_vocab_size)
        self.fc2 = nn.Linear(ans_vocab_size, ans_vocab_size)
        self.embed_size = embed_size
        self.qst_vocab_size = qst_vocab_size
        self.ans_vocab_size = ans_vocab_size
        self.word_embed_size = word_embed_size
        self.num_layers = num_layers
        self.hidden_size = hidden_size
        self.criterion = nn.CrossEntropyLoss()

    def forward(self, img, qst):
        z = torch.mul(self.img_encoder(img), self.qst_encoder(qst))
        z = self.dropout(self.tanh(z))
        z = self.dropout(self.tanh(self.fc1(z)))
        return self.fc2(z)

    def new(self):
        twin = VqaModel(self.embed_size, self.qst_vocab_size, self.ans_vocab_size, self.word_embed_size,
                        self.num_layers, self.hidden_size, self.pretrained)
        _copy_dropout(self, twin)
        return twin.to(config.DEVICE)

    def _loss(self, images, questions, labels):
        return self.criterion(self(images, questions), labels)

    def _soft_loss(self, images, questions, labels, pseudo_qst, pseudo_labels):
        loss_1 = self.criterion(self(images, questions), labels)
        loss_2 = softXEnt(self(images, pseudo_qst), pseudo_labels)
        return loss_1 + config.W_LAMBDA * loss_2
