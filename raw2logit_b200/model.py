"""Lightning-free equivalent of the reference's training harness (reference model.py:15-151), the caller of the ISP
hot path: ``LitModel.forward = classifier(augmentation(processor(x)))`` with the same constructor arguments, freeze
rules, ``adv_parameters`` substring selection and ``train()`` override.  pytorch_lightning / mlflow are not part of
this image, so logging and the Trainer loop are out; ``update_step`` returns the loss and
``configure_optimizers`` builds the same Adam.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def resnet_model(model='resnet18', pretrained=False, in_channels=3, fc_out_features=2):
    """torchvision ResNet with a fresh ``fc`` head (reference model.py:15-23; weights are random here: no network)."""
    import torchvision.models as tvm
    name = model.lower()
    if name not in ('resnet18', 'resnet34', 'resnet50'):
        raise ValueError(f"unknown classifier {model}")
    net = getattr(tvm, name)(weights=None)
    net.fc = nn.Linear(in_features=net.fc.in_features, out_features=fc_out_features, bias=True)
    return net


def _freeze(module):
    """pl.LightningModule.freeze: no gradients, eval mode (reference model.py:64-68)."""
    for p in module.parameters():
        p.requires_grad = False
    module.eval()


class LitModel(nn.Module):
    def __init__(self, classifier, loss, lr=1e-3, weight_decay=0, loss_aux=None, adv_training=False,
                 adv_parameters='all', metrics=None, processor=None, augmentation=None, is_segmentation_task=False,
                 augmentation_on_eval=False, metrics_on_training=True, freeze_classifier=False,
                 freeze_processor=False):
        super().__init__()
        self.classifier = classifier
        self.processor = processor
        self.lr = lr
        self.weight_decay = weight_decay
        self.loss_fn = loss
        self.loss_aux_fn = loss_aux
        self.adv_training = adv_training
        self.metrics = metrics
        self.augmentation = augmentation
        self.is_segmentation_task = is_segmentation_task
        self.augmentation_on_eval = augmentation_on_eval
        self.metrics_on_training = metrics_on_training
        self.freeze_classifier = freeze_classifier
        self.freeze_processor = freeze_processor

        for p in self.parameters():                                   # unfreeze() (model.py:64)
            p.requires_grad = True
        if freeze_classifier:
            _freeze(self.classifier)
        if freeze_processor:
            _freeze(self.processor)
        if adv_training and adv_parameters != 'all':                  # model.py:70-75: one group by name substring
            _freeze(self.processor)
            for name, p in self.processor.named_parameters():
                if adv_parameters in name:
                    p.requires_grad = True
        self.train(self.training)

    def forward(self, x):
        x = self.processor(x)
        if self.augmentation is not None and (self.training or self.augmentation_on_eval):
            x = self.augmentation(x, retain_state=self.is_segmentation_task)
        return self.classifier(x)

    def update_step(self, batch, step_name='train'):
        x, y = batch
        logits = self(x)
        if self.augmentation is not None and self.is_segmentation_task and (self.training or self.augmentation_on_eval):
            y = self.augmentation(y, mask_transform=True).contiguous()
        loss = self.loss_fn(logits, y)
        if self.loss_aux_fn is not None:
            loss = loss + self.loss_aux_fn(x)
        return loss

    def training_step(self, batch, batch_idx=0):
        return self.update_step(batch, 'train')

    def validation_step(self, batch, batch_idx=0):
        return self.update_step(batch, 'val')

    def train(self, mode=True):
        self.training = mode
        # no BatchNorm updates in adversarial training or with a frozen processor (model.py:136-142)
        if self.processor is not None:
            self.processor.train(mode=mode and not self.freeze_processor and not self.adv_training)
        self.classifier.train(mode=mode and not self.freeze_classifier)
        return self

    def configure_optimizers(self):
        self.optimizer = torch.optim.Adam(self.parameters(), self.lr, weight_decay=self.weight_decay)
        return self.optimizer


class SmallUNet(nn.Module):
    """Plain-PyTorch stand-in for the reference's smp.UnetPlusPlus segmentation model (train.py:218;
    segmentation_models_pytorch is not installable here): 3 -> 1 logits, H and W multiples of 8."""

    def __init__(self, width=32):
        super().__init__()
        def block(i, o):
            return nn.Sequential(nn.Conv2d(i, o, 3, padding=1, bias=False), nn.BatchNorm2d(o), nn.ReLU(inplace=True),
                                 nn.Conv2d(o, o, 3, padding=1, bias=False), nn.BatchNorm2d(o), nn.ReLU(inplace=True))
        w = width
        self.e1, self.e2, self.e3, self.mid = block(3, w), block(w, 2 * w), block(2 * w, 4 * w), block(4 * w, 8 * w)
        self.d3, self.d2, self.d1 = block(12 * w, 4 * w), block(6 * w, 2 * w), block(3 * w, w)
        self.head = nn.Conv2d(w, 1, 1)

    def forward(self, x):
        e1 = self.e1(x)
        e2 = self.e2(F.max_pool2d(e1, 2))
        e3 = self.e3(F.max_pool2d(e2, 2))
        m = self.mid(F.max_pool2d(e3, 2))
        d3 = self.d3(torch.cat([F.interpolate(m, scale_factor=2.0), e3], 1))
        d2 = self.d2(torch.cat([F.interpolate(d3, scale_factor=2.0), e2], 1))
        d1 = self.d1(torch.cat([F.interpolate(d2, scale_factor=2.0), e1], 1))
        return self.head(d1)


def dice_loss(logits, target, eps=1.0):
    """Binary Dice loss on logits (the role of smp.losses.DiceLoss(mode='binary', from_logits=True), train.py:236)."""
    p = torch.sigmoid(logits).flatten(1)
    t = target.reshape(p.shape).to(p.dtype)
    inter = (p * t).sum(1)
    return (1 - (2 * inter + eps) / (p.sum(1) + t.sum(1) + eps)).mean()
