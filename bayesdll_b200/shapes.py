"""Parameter shapes of the backbones the BASELINE configs name (networks/__init__.py:15-54 of the reference):
the backbones themselves stay torchvision / plain PyTorch; this only builds them (on the meta device when only
the shapes are needed) the way the reference's ``create_backbone`` does, incl. ``net.readout_name``."""
import torch
import torch.nn as nn


class MLP(nn.Module):
    """784-1000-1000-1000-10 MLP (networks/small_nets.py:7-45): ``layers`` Sequential + ``classifier``."""

    def __init__(self, input_dim=784, output_dim=10, width=1000, depth=3):
        super().__init__()
        self.input_dim = input_dim
        layers, hin = [], input_dim
        for _ in range(depth):
            layers += [nn.Linear(hin, width), nn.ReLU()]
            hin = width
        self.layers = nn.Sequential(*layers)
        self.classifier = nn.Linear(width, output_dim)

    def forward(self, x):
        return self.classifier(self.layers(x.view(-1, self.input_dim)))


def create_backbone(name, num_classes=37):
    if name == "mlp_mnist":
        net = MLP()
        net.readout_name = "classifier"
        return net
    import torchvision
    if name == "resnet101":
        net = torchvision.models.resnet101()
        net.fc = nn.Linear(2048, num_classes)
        head, net.readout_name = net.fc, "fc"
    elif name == "vit_l_32":
        net = torchvision.models.vit_l_32()
        net.heads.head = nn.Linear(1024, num_classes)
        head, net.readout_name = net.heads.head, "heads.head"
    else:
        raise NotImplementedError(name)
    for p in head.parameters():
        if p.dim() > 1:
            nn.init.kaiming_normal_(p, nonlinearity="relu")
        else:
            nn.init.zeros_(p)
    return net


def named_shapes(name, num_classes=37):
    """[(param_name, shape)] in named_parameters() order plus readout_name, without allocating weights."""
    with torch.device("meta"):
        net = create_backbone(name, num_classes)
    return [(n, tuple(p.shape)) for n, p in net.named_parameters()], net.readout_name
