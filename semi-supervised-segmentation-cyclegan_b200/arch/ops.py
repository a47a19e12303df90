"""Mirror of the hot part of the reference's arch/ops.py (reference arch/ops.py:7-80): same
function names, signatures, defaults, error strings and module trees, so `state_dict()` keys and
the N(0, 0.02) initialisation order are identical.  The modules built here are parameter
containers + a stock-torch CPU path (the reference's own `gpu_ids=[]` behaviour); on CUDA the
owning network (generators.ResnetGenerator / discriminators.NLayerDiscriminator) executes them as
fused stages through libsscg_b200.so.
"""
import functools

import torch
import torch.nn as nn
from torch.nn import init


def get_norm_layer(norm_type='instance'):
    # reference arch/ops.py:7-14
    if norm_type == 'batch':
        norm_layer = functools.partial(nn.BatchNorm2d, affine=True)
    elif norm_type == 'instance':
        norm_layer = functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    else:
        raise NotImplementedError('normalization layer [%s] is not found' % norm_type)
    return norm_layer


def init_weights(net, init_type='normal', gain=0.02):
    # reference arch/ops.py:16-28 — Conv*/Linear weight ~ N(0, gain), bias = 0; BatchNorm2d weight ~ N(1, gain).
    # Draws from the global torch RNG in module-traversal order, exactly like the reference.
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, 'weight') and (classname.find('Conv') != -1 or classname.find('Linear') != -1):
            init.normal_(m.weight.data, 0.0, gain)
            if hasattr(m, 'bias') and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find('BatchNorm2d') != -1:
            init.normal_(m.weight.data, 1.0, gain)
            init.constant_(m.bias.data, 0.0)

    print('Network initialized with weights sampled from N(0,0.02).')
    net.apply(init_func)


def init_network(net, gpu_ids=[]):
    # reference arch/ops.py:31-37 — move to gpu_ids[0] first, then initialise (so a GPU-built net
    # draws its weights from the CUDA generator)
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
        net.cuda(gpu_ids[0])
    init_weights(net)
    return net


def conv_norm_lrelu(in_dim, out_dim, kernel_size, stride=1, padding=0, norm_layer=nn.BatchNorm2d, bias=False):
    # reference arch/ops.py:40-44
    return nn.Sequential(
        nn.Conv2d(in_dim, out_dim, kernel_size, stride, padding, bias=bias),
        norm_layer(out_dim), nn.LeakyReLU(0.2, True))


def conv_norm_relu(in_dim, out_dim, kernel_size, stride=1, padding=0, norm_layer=nn.BatchNorm2d, bias=False):
    # reference arch/ops.py:46-50
    return nn.Sequential(
        nn.Conv2d(in_dim, out_dim, kernel_size, stride, padding, bias=bias),
        norm_layer(out_dim), nn.ReLU(True))


def dconv_norm_relu(in_dim, out_dim, kernel_size, stride=1, padding=0, output_padding=0, norm_layer=nn.BatchNorm2d,
                    bias=False):
    # reference arch/ops.py:52-57
    return nn.Sequential(
        nn.ConvTranspose2d(in_dim, out_dim, kernel_size, stride, padding, output_padding, bias=bias),
        norm_layer(out_dim), nn.ReLU(True))


class ResidualBlock(nn.Module):
    # reference arch/ops.py:59-74
    def __init__(self, dim, norm_layer, use_dropout, use_bias):
        super(ResidualBlock, self).__init__()
        res_block = [nn.ReflectionPad2d(1),
                     conv_norm_relu(dim, dim, kernel_size=3, norm_layer=norm_layer, bias=use_bias)]
        if use_dropout:
            res_block += [nn.Dropout(0.5)]
        res_block += [nn.ReflectionPad2d(1),
                      nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias),
                      norm_layer(dim)]
        self.res_block = nn.Sequential(*res_block)
        self.use_dropout = use_dropout

    def forward(self, x):
        return x + self.res_block(x)


def set_grad(nets, requires_grad=False):
    # reference arch/ops.py:77-80
    for net in nets:
        for param in net.parameters():
            param.requires_grad = requires_grad


def _is_instance_norm(norm_layer):
    if type(norm_layer) == functools.partial:
        return norm_layer.func == nn.InstanceNorm2d
    return norm_layer == nn.InstanceNorm2d
