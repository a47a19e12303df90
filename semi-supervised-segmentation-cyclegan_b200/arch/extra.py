"""Generator / discriminator families of the reference that are NOT on the B200 hot path, as stock torch.nn modules
with the reference's module tree (so `state_dict()` keys, shapes and checkpoints interoperate) — SURVEY.md §7.2
("non-hot names route to stock torch") and §8 f5:

  * DeepLab-v2 / ResNet-101 (`netG='deeplab'`, reference arch/generators.py:320-441,510-511): what HEAD's model.py
    actually instantiates for Gis / Gsi (model.py:215-219); frozen BatchNorm, dilated layer3 / layer4, a four-branch
    atrous classifier whose forward returns after the second branch (generators.py:378-382 — kept as is: it is the
    reference's behaviour and checkpoints depend on it);
  * U-Net (`unet_128`, `unet_256`, generators.py:7-63);
  * FCDiscriminator (`fc_disc`, discriminators.py:8-39).

ENet / LEDNet (`enet`, `lednet_*`, generators.py:98-318 + 500 lines of ops.py) are delegated to the reference's own
`arch` package when it is importable next to this one (a user who swaps `from arch import define_Gen` for
`from sscg_b200.arch import define_Gen` still has it) and raise NotImplementedError otherwise.  None of this runs in
this repo's kernels: cuDNN / ATen execute it, exactly as in the reference.
"""
import functools
import importlib
import sys

import torch
import torch.nn as nn


# ---------------------------------------------------------------------------------------------------------------
# DeepLab-v2 (ResNet-101 backbone), reference arch/generators.py:320-441
# ---------------------------------------------------------------------------------------------------------------
def _frozen_bn(c):
    bn = nn.BatchNorm2d(c)
    for p in bn.parameters():            # generators.py:328-329,335-336,339-340,390-391,418-419
        p.requires_grad = False
    return bn


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False)
        self.bn1 = _frozen_bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=dilation, bias=False, dilation=dilation)
        self.bn2 = _frozen_bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = _frozen_bn(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        out += x if self.downsample is None else self.downsample(x)
        return self.relu(out)


class Classifier_Module(nn.Module):
    def __init__(self, dilation_series, padding_series, num_classes):
        super().__init__()
        self.conv2d_list = nn.ModuleList(
            nn.Conv2d(2048, num_classes, kernel_size=3, stride=1, padding=p, dilation=d, bias=True)
            for d, p in zip(dilation_series, padding_series))
        for m in self.conv2d_list:
            m.weight.data.normal_(0, 0.01)

    def forward(self, x):
        out = self.conv2d_list[0](x)
        for i in range(len(self.conv2d_list) - 1):
            out += self.conv2d_list[i + 1](x)
            return out          # sic: the reference returns inside the loop (generators.py:380-382): branches 0 + 1 only


class DeepLabResNet(nn.Module):
    """`ResNet(in_channels, Bottleneck, [3, 4, 23, 3], num_classes)` of the reference (generators.py:384-441)."""

    def __init__(self, in_channels, block, layers, num_classes):
        self.inplanes = 64
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = _frozen_bn(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=True)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=4)
        self.layer5 = Classifier_Module([6, 12, 18, 24], [6, 12, 18, 24], num_classes)
        for m in self.modules():         # generators.py:400-406 (init_network re-initialises the convs afterwards)
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, 0.01)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion or dilation in (2, 4):
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                _frozen_bn(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, dilation=dilation, downsample=downsample)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes, dilation=dilation) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        return self.layer5(self.layer4(self.layer3(self.layer2(self.layer1(x)))))

    # optimizer helpers of the reference (generators.py:443-485)
    def get_1x_lr_params_NOscale(self):
        for part in (self.conv1, self.bn1, self.layer1, self.layer2, self.layer3, self.layer4):
            for mod in part.modules():
                for p in mod.parameters(recurse=False):
                    if p.requires_grad:
                        yield p

    def get_10x_lr_params(self):
        yield from self.layer5.parameters()

    def optim_parameters(self, args):
        return [{"params": self.get_1x_lr_params_NOscale(), "lr": args.learning_rate},
                {"params": self.get_10x_lr_params(), "lr": 10 * args.learning_rate}]


def deeplab(input_nc, output_nc):
    return DeepLabResNet(input_nc, Bottleneck, [3, 4, 23, 3], output_nc)


# ---------------------------------------------------------------------------------------------------------------
# U-Net, reference arch/generators.py:7-63
# ---------------------------------------------------------------------------------------------------------------
class UnetSkipConnectionBlock(nn.Module):
    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        self.outermost = outermost
        func = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        use_bias = func == nn.InstanceNorm2d
        input_nc = outer_nc if input_nc is None else input_nc
        downconv = nn.Conv2d(input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        if outermost:
            model = [downconv, submodule, nn.ReLU(True),
                     nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)]
        elif innermost:
            model = [nn.LeakyReLU(0.2, True), downconv, nn.ReLU(True),
                     nn.ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias),
                     norm_layer(outer_nc)]
        else:
            model = [nn.LeakyReLU(0.2, True), downconv, norm_layer(inner_nc), submodule, nn.ReLU(True),
                     nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias),
                     norm_layer(outer_nc)]
            if use_dropout:
                model.append(nn.Dropout(0.5))
        self.model = nn.Sequential(*model)

    def forward(self, x):
        return self.model(x) if self.outermost else torch.cat([x, self.model(x)], 1)


class UnetGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        blk = UnetSkipConnectionBlock(ngf * 8, ngf * 8, submodule=None, norm_layer=norm_layer, innermost=True)
        for _ in range(num_downs - 5):
            blk = UnetSkipConnectionBlock(ngf * 8, ngf * 8, submodule=blk, norm_layer=norm_layer, use_dropout=use_dropout)
        for outer, inner in ((ngf * 4, ngf * 8), (ngf * 2, ngf * 4), (ngf, ngf * 2)):
            blk = UnetSkipConnectionBlock(outer, inner, submodule=blk, norm_layer=norm_layer)
        self.unet_model = UnetSkipConnectionBlock(output_nc, ngf, input_nc=input_nc, submodule=blk, outermost=True,
                                                  norm_layer=norm_layer)

    def forward(self, input):
        return self.unet_model(input)


# ---------------------------------------------------------------------------------------------------------------
# FCDiscriminator (AdvSemiSeg), reference arch/discriminators.py:8-39
# ---------------------------------------------------------------------------------------------------------------
class FCDiscriminator(nn.Module):
    def __init__(self, num_classes, ndf=64):
        super().__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(ndf, ndf * 2, kernel_size=4, stride=2, padding=1)
        self.conv3 = nn.Conv2d(ndf * 2, ndf * 4, kernel_size=4, stride=2, padding=1)
        self.conv4 = nn.Conv2d(ndf * 4, ndf * 8, kernel_size=4, stride=2, padding=1)
        self.classifier = nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def forward(self, x):
        for conv in (self.conv1, self.conv2, self.conv3, self.conv4):
            x = self.leaky_relu(conv(x))
        return self.classifier(x)


# ---------------------------------------------------------------------------------------------------------------
# ENet / LEDNet: the reference's own modules, when its `arch` package is importable
# ---------------------------------------------------------------------------------------------------------------
def reference_generators():
    """The reference's arch.generators module, or None.  (Not this package: a module named `arch` that defines ENet.)"""
    for name in ("arch.generators",):
        try:
            mod = sys.modules.get(name) or importlib.import_module(name)
        except Exception:                # noqa: BLE001 — absent or broken: not available
            continue
        if hasattr(mod, "ENet") and hasattr(mod, "LEDNet") and not mod.__name__.startswith("sscg_b200"):
            return mod
    return None
