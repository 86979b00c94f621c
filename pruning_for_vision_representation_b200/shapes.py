"""Element counts of the prunable tensors (every nn.Conv2d / nn.Linear weight, named_modules()
order — train.py:263-265, 333-336) of the models BASELINE.json names, without building them."""


def _resnet(block, layers, num_classes=1000):
    out = [64 * 3 * 7 * 7]
    inplanes = 64
    exp = 4 if block == "bottleneck" else 1
    for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
        stride = 1 if i == 0 else 2
        for b in range(n):
            if block == "bottleneck":
                out += [planes * inplanes, planes * planes * 9, planes * 4 * planes]
            else:
                out += [planes * inplanes * 9, planes * planes * 9]
            if b == 0 and (stride != 1 or inplanes != planes * exp):
                out.append(planes * exp * inplanes)      # downsample conv comes after the block's convs
            inplanes = planes * exp
    out.append(num_classes * 512 * exp)
    return out


def _vit(patch, hidden, mlp, layers, num_classes=1000):
    out = [hidden * 3 * patch * patch]
    for _ in range(layers):
        # nn.MultiheadAttention.in_proj_weight is a bare Parameter (not pruned); out_proj is a Linear
        out += [hidden * hidden, mlp * hidden, hidden * mlp]
    out.append(num_classes * hidden)
    return out


_MODELS = {
    "resnet18": lambda: _resnet("basic", (2, 2, 2, 2)),
    "resnet50": lambda: _resnet("bottleneck", (3, 4, 6, 3)),
    "resnet152": lambda: _resnet("bottleneck", (3, 8, 36, 3)),
    "vit_b_16": lambda: _vit(16, 768, 3072, 12),
    "vit_b_32": lambda: _vit(32, 768, 3072, 12),
    "vit_l_16": lambda: _vit(16, 1024, 4096, 24),
}


def prunable_numels(model_name):
    return list(_MODELS[model_name]())


def model_names():
    return sorted(_MODELS)
