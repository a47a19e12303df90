"""Mirror of the hot part of the reference's arch/generators.py: ResnetGenerator
(reference arch/generators.py:65-95) and the define_Gen factory (reference arch/generators.py:487-515).

The module tree, parameter names and shapes are identical to the reference (so reference
checkpoints load unchanged and vice versa).  On CUDA with instance norm the forward
runs as fused stages through libsscg_b200.so (see engine.py); CPU tensors take the stock-torch
module path, which is the reference's own `gpu_ids=[]` behaviour.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from ..engine import StageSpec
from ..runtime import NetRunner
from .ops import (ResidualBlock, _is_instance_norm, conv_norm_relu, dconv_norm_relu, get_norm_layer, init_network)


class ResnetGenerator(nn.Module):
    def __init__(self, input_nc=3, output_nc=3, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=True, num_blocks=6,
                 softmax=False):
        super(ResnetGenerator, self).__init__()
        use_bias = _is_instance_norm(norm_layer)
        res_model = [nn.ReflectionPad2d(3),
                     conv_norm_relu(input_nc, ngf * 1, 7, norm_layer=norm_layer, bias=use_bias),
                     conv_norm_relu(ngf * 1, ngf * 2, 3, 2, 1, norm_layer=norm_layer, bias=use_bias),
                     conv_norm_relu(ngf * 2, ngf * 4, 3, 2, 1, norm_layer=norm_layer, bias=use_bias)]
        for _ in range(num_blocks):
            res_model += [ResidualBlock(ngf * 4, norm_layer, use_dropout, use_bias)]
        res_model += [dconv_norm_relu(ngf * 4, ngf * 2, 3, 2, 1, 1, norm_layer=norm_layer, bias=use_bias),
                      dconv_norm_relu(ngf * 2, ngf * 1, 3, 2, 1, 1, norm_layer=norm_layer, bias=use_bias),
                      nn.ReflectionPad2d(3),
                      nn.Conv2d(ngf, output_nc, 7)]
        if not softmax:   # the `*_softmax` variants emit raw logits (reference generators.py:81-85)
            res_model += [nn.Tanh()]
        self.res_model = nn.Sequential(*res_model)

        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        self.num_blocks, self.use_dropout, self.softmax = num_blocks, use_dropout, softmax
        self.fusable = use_bias          # instance norm (affine=False): the configuration the kernels implement
        self.precision = None            # None -> runtime.default_precision()
        self._runner = NetRunner(self._stage_specs, _residual_plan) if self.fusable else None

    # ---- stage list for the engine ------------------------------------------------------------
    def _stage_specs(self):
        m = self.res_model
        nb, ngf = self.num_blocks, self.ngf
        specs = [
            StageSpec("window", 7, 1, 3, 3, True, self.input_nc, ngf, True, L.ACT_RELU, m[1][0].weight, m[1][0].bias,
                      name="stem"),
            StageSpec("conv", 3, 2, 1, 0, False, ngf, ngf * 2, True, L.ACT_RELU, m[2][0].weight, m[2][0].bias,
                      name="down1"),
            StageSpec("conv", 3, 2, 1, 0, False, ngf * 2, ngf * 4, True, L.ACT_RELU, m[3][0].weight, m[3][0].bias,
                      name="down2"),
        ]
        dim = ngf * 4
        for b in range(nb):
            blk = m[4 + b].res_block
            c1 = blk[1][0]
            c2 = blk[4] if self.use_dropout else blk[3]
            specs.append(StageSpec("conv", 3, 1, 1, 1, True, dim, dim, True, L.ACT_RELU, c1.weight, c1.bias,
                                   dropout=self.use_dropout, name="res%d.conv1" % b))
            specs.append(StageSpec("conv", 3, 1, 1, 1, True, dim, dim, True, L.ACT_NONE, c2.weight, c2.bias,
                                   residual_from=3 + 2 * b, name="res%d.conv2" % b))
        j = 4 + nb
        specs.append(StageSpec("convT", 3, 2, 1, 0, False, dim, ngf * 2, True, L.ACT_RELU, m[j][0].weight, m[j][0].bias,
                               name="up1"))
        specs.append(StageSpec("convT", 3, 2, 1, 0, False, ngf * 2, ngf, True, L.ACT_RELU, m[j + 1][0].weight,
                               m[j + 1][0].bias, name="up2"))
        head = m[j + 3]
        specs.append(StageSpec("conv", 7, 1, 3, 3, True, ngf, self.output_nc, False,
                               L.ACT_NONE if self.softmax else L.ACT_TANH, head.weight, head.bias, final=True,
                               name="head"))
        return specs

    def forward(self, x):
        if x.is_cuda and self.fusable:
            return self._runner(x, self.training, self.use_dropout, self.precision)
        return self.res_model(x)

    def forward_parts(self, parts):
        """ONE batched pass over several inputs (float N x C x H x W tensors and / or int64 label maps N x 1 x H x W that
        stand for their one-hot encodings): returns the concatenation of what separate calls would return.  The
        normalisation is per sample, so batching changes no result; it halves the launches and lengthens every
        kernel's work list (step.py uses it for the two passes that share a network and have independent inputs)."""
        if parts[0].is_cuda and self.fusable:
            return self._runner([p.long() if not p.is_floating_point() else p for p in parts], self.training, self.use_dropout,
                                self.precision)
        dense = []
        for p in parts:
            if not p.is_floating_point():
                p = torch.zeros(p.size(0), self.input_nc, p.size(2), p.size(3), dtype=torch.float32,
                                device=p.device).scatter_(1, p.long(), 1)
            dense.append(p.float())
        return self.forward(torch.cat(dense))

    def forward_onehot(self, labels):
        """forward(make_one_hot(labels, input_nc)) for an int64 label map N x 1 x H x W (model.py:385) without
        materialising the one-hot tensor on the fused path."""
        if labels.is_cuda and self.fusable:
            return self._runner(labels.long(), self.training, self.use_dropout, self.precision)
        one_hot = torch.zeros(labels.size(0), self.input_nc, labels.size(2), labels.size(3), dtype=torch.float32,
                              device=labels.device).scatter_(1, labels.long(), 1)
        return self.forward(one_hot)


def _residual_plan(specs):
    """Backward routing of the residual skip connections (x + block(x), reference ops.py:73-74).

    T_j = total gradient w.r.t. the output P_j of residual block j-1 (= input of block j).
    Stage conv2 of block b produces P_{b+1}; its incoming gradient is fold(dgrad of block b+1's
    conv1) + T_{b+2}.  Returns {stage index: (skip tag, g_out tag)} for engine.NetPlan."""
    idx = [i for i, s in enumerate(specs) if s.residual_from is not None]
    B = len(idx)
    if B == 0:
        return {}
    first = idx[0] - 1             # conv1 of block 0 == act index of P_0
    up1 = first + 2 * B            # stage that consumes P_B; gact[up1] is T_B

    def tag(j):
        return ("g", up1) if j == B else ("t", j % 2)

    plan = {}
    for b in range(B):
        stage = first + 2 * b + 1
        skip = tag(b + 2) if b + 2 <= B else None
        gout = tag(b + 1) if b + 1 < B else None
        plan[stage] = (skip, gout)
    plan[first - 1] = (tag(1), None)   # the stage that produces P_0 (down2)
    return plan


def define_Gen(input_nc, output_nc, ngf, netG, norm='batch', use_dropout=False, gpu_ids=[0]):
    # reference arch/generators.py:487-515
    gen_net = None
    norm_layer = get_norm_layer(norm_type=norm)

    if netG == 'resnet_9blocks':
        gen_net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                                  num_blocks=9, softmax=False)
    elif netG == 'resnet_9blocks_softmax':
        gen_net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                                  num_blocks=9, softmax=True)
    elif netG == 'resnet_6blocks':
        gen_net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                                  num_blocks=6, softmax=False)
    elif netG == 'resnet_6blocks_softmax':
        gen_net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                                  num_blocks=6, softmax=True)
    elif netG in ('unet_128', 'unet_256', 'deeplab'):
        # not on the B200 hot path: stock torch.nn modules with the reference's module tree (arch/extra.py; SURVEY.md
        # §7.2, §8 f5) — cuDNN / ATen execute them, exactly as in the reference (generators.py:499-502,510-511)
        from . import extra
        if netG == 'deeplab':
            gen_net = extra.deeplab(input_nc, output_nc)
        else:
            gen_net = extra.UnetGenerator(input_nc, output_nc, 7 if netG == 'unet_128' else 8, ngf,
                                          norm_layer=norm_layer, use_dropout=use_dropout)
    elif netG in ('enet', 'lednet_128', 'lednet_256'):
        # reference generators.py:503-509 — delegated to the reference's own modules when its `arch` package is
        # importable next to this one (~700 lines of blocks outside the hot path are not restated here)
        from . import extra
        ref = extra.reference_generators()
        if ref is None:
            raise NotImplementedError('Generator model name [%s] is outside the B200 hot path and the reference '
                                      'package `arch` (which defines it) is not importable' % netG)
        if netG == 'enet':
            gen_net = ref.ENet(num_classes=output_nc, encoder_relu=False, decoder_relu=True)
        else:
            gen_net = ref.LEDNet(in_channels=input_nc, n_classes=output_nc, encoder_relu=False, decoder_relu=True,
                                 image_dim=128 if netG == 'lednet_128' else 256)
    else:
        raise NotImplementedError('Generator model name [%s] is not recognized' % netG)

    return init_network(gen_net, gpu_ids)
