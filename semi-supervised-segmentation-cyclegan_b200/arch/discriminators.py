"""Mirror of the hot part of the reference's arch/discriminators.py: the 70x70 PatchGAN
NLayerDiscriminator (reference arch/discriminators.py:42-63) and the define_Dis factory
(reference arch/discriminators.py:84-101).  `pixel` (discriminators.py:66-80) is kept routable as a
small stock-torch module because HEAD's training loop instantiates it (model.py:220-222,229).
"""
import torch
import torch.nn as nn

from .. import _lib as L
from ..engine import StageSpec
from ..runtime import NetRunner
from .ops import _is_instance_norm, conv_norm_lrelu, get_norm_layer, init_network


class NLayerDiscriminator(nn.Module):
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_bias=False):
        super(NLayerDiscriminator, self).__init__()
        dis_model = [nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=1),
                     nn.LeakyReLU(0.2, True)]
        nf_mult = 1
        nf_mult_prev = 1
        chans = []
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** n, 8)
            dis_model += [conv_norm_lrelu(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=4, stride=2,
                                          norm_layer=norm_layer, padding=1, bias=use_bias)]
            chans.append((ndf * nf_mult_prev, ndf * nf_mult, 2))
        nf_mult_prev = nf_mult
        nf_mult = min(2 ** n_layers, 8)
        dis_model += [conv_norm_lrelu(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=4, stride=1,
                                      norm_layer=norm_layer, padding=1, bias=use_bias)]
        chans.append((ndf * nf_mult_prev, ndf * nf_mult, 1))
        dis_model += [nn.Conv2d(ndf * nf_mult, 1, kernel_size=4, stride=1, padding=1)]
        self.dis_model = nn.Sequential(*dis_model)

        self.input_nc, self.ndf, self.n_layers = input_nc, ndf, n_layers
        self._chans = chans
        self.fusable = _is_instance_norm(norm_layer) and use_bias
        self.precision = None
        self._runner = NetRunner(self._stage_specs) if self.fusable else None

    def _stage_specs(self):
        m = self.dis_model
        specs = [StageSpec("window", 4, 2, 1, 1, False, self.input_nc, self.ndf, False, L.ACT_LRELU, m[0].weight,
                           m[0].bias, name="d0")]
        idx = 2
        for (ci, co, st) in self._chans:
            conv = m[idx][0]
            specs.append(StageSpec("conv", 4, st, 1, 0, False, ci, co, True, L.ACT_LRELU, conv.weight, conv.bias,
                                   name="d%d" % (idx - 1)))
            idx += 1
        tail = m[idx]
        specs.append(StageSpec("conv", 4, 1, 1, 0, False, self._chans[-1][1], 1, False, L.ACT_NONE, tail.weight,
                               tail.bias, final=True, name="dtail"))
        return specs

    def forward(self, input):
        if input.is_cuda and self.fusable:
            return self._runner(input, self.training, False, self.precision)
        return self.dis_model(input)

    def forward_parts(self, parts):
        """ONE batched pass over several inputs (float N x C x H x W tensors and / or int64 label maps N x 1 x H x W that
        stand for their one-hot encodings): returns the concatenation of what separate calls would return.  The
        normalisation is per sample, so batching changes no result; it halves the launches and lengthens every
        kernel's work list (step.py uses it for the two passes that share a network and have independent inputs)."""
        if parts[0].is_cuda and self.fusable:
            return self._runner([p.long() if not p.is_floating_point() else p for p in parts], self.training, False,
                                self.precision)
        dense = []
        for p in parts:
            if not p.is_floating_point():
                p = torch.zeros(p.size(0), self.input_nc, p.size(2), p.size(3), dtype=torch.float32,
                                device=p.device).scatter_(1, p.long(), 1)
            dense.append(p.float())
        return self.forward(torch.cat(dense))

    def forward_onehot(self, labels):
        """forward(make_one_hot(labels, input_nc)) for an int64 label map N x 1 x H x W (model.py:435-438,506-512)
        without materialising the one-hot tensor on the fused path."""
        if labels.is_cuda and self.fusable:
            return self._runner(labels.long(), self.training, False, self.precision)
        one_hot = torch.zeros(labels.size(0), self.input_nc, labels.size(2), labels.size(3), dtype=torch.float32,
                              device=labels.device).scatter_(1, labels.long(), 1)
        return self.forward(one_hot)


class PixelDiscriminator(nn.Module):
    # reference arch/discriminators.py:66-80 — 1x1 convolutions; stock torch (not on the hot path)
    def __init__(self, input_nc, ndf=64, norm_layer=nn.BatchNorm2d, use_bias=False):
        super(PixelDiscriminator, self).__init__()
        self.dis_model = nn.Sequential(
            nn.Conv2d(input_nc, ndf, kernel_size=1, stride=1, padding=0),
            nn.LeakyReLU(0.2, True),
            nn.Conv2d(ndf, ndf * 2, kernel_size=1, stride=1, padding=0, bias=use_bias),
            norm_layer(ndf * 2),
            nn.LeakyReLU(0.2, True),
            nn.Conv2d(ndf * 2, 1, kernel_size=1, stride=1, padding=0, bias=use_bias))

    def forward(self, input):
        return self.dis_model(input)


def define_Dis(input_nc, ndf, netD, n_layers_D=3, norm='batch', gpu_ids=[0]):
    # reference arch/discriminators.py:84-101
    dis_net = None
    norm_layer = get_norm_layer(norm_type=norm)
    use_bias = _is_instance_norm(norm_layer)

    if netD == 'n_layers':
        dis_net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer=norm_layer, use_bias=use_bias)
    elif netD == 'pixel':
        dis_net = PixelDiscriminator(input_nc, ndf, norm_layer=norm_layer, use_bias=use_bias)
    elif netD == 'fc_disc':
        # AdvSemiSeg FCDiscriminator (reference discriminators.py:8-39,95-96): stock torch, outside the hot path
        from .extra import FCDiscriminator
        dis_net = FCDiscriminator(input_nc, ndf)
    else:
        raise NotImplementedError('Discriminator model name [%s] is not recognized' % netD)

    return init_network(dis_net, gpu_ids)
