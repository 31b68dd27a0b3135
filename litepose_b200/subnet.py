"""Sub-network extraction from a (super)network checkpoint.

The reference trains one supernet and obtains LitePose-XS...L by weight transfer (weight_transfer.py:75-146): every
tensor of the sub-network is the leading slice of the same-named supernet tensor - conv [:out, :in], transposed conv
[:in, :out], depthwise [:mid], BatchNorm vectors [:n] - and blocks beyond the sub-network's depth are dropped.  Because
all of those are prefix slices in every dimension, one rule covers them."""
import torch


def extract_subnet_state_dict(super_state_dict, sub_state_dict_like):
    """super_state_dict: checkpoint of the larger network (keys optionally prefixed '1.' by network_to_half);
    sub_state_dict_like: state_dict (or {name: shape}) of the target architecture.  Returns a state_dict for the target:
    floating-point tensors as float32 clones (weight_transfer.py does .float().clone()), counters unchanged."""
    sup = {(k[2:] if k.startswith("1.") else k): v for k, v in super_state_dict.items()}
    out = {}
    for name, like in sub_state_dict_like.items():
        shape = tuple(like.shape) if hasattr(like, "shape") else tuple(like)
        if name not in sup:
            raise KeyError("supernet checkpoint has no tensor %r" % name)
        src = sup[name]
        if src.dim() != len(shape) or any(s < d for s, d in zip(src.shape, shape)):
            raise ValueError("%s: supernet tensor %s cannot provide %s" % (name, tuple(src.shape), shape))
        piece = src[tuple(slice(0, d) for d in shape)]
        out[name] = piece.float().clone() if piece.is_floating_point() else piece.clone()
    return out


def extract_subnet(super_state_dict, cfg, arch):
    """Builds the target LitePose (litepose_b200 drop-in module) for ``arch`` and loads the extracted weights."""
    from litepose_b200.lib.models.pose_mobilenet import get_pose_net
    net = get_pose_net(cfg, False, arch)
    net.load_state_dict(extract_subnet_state_dict(super_state_dict, net.state_dict()), strict=True)
    return net.eval()
