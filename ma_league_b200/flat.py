"""Flat fp32 parameter buffers behind ordinary nn.Parameters.

The kernels take ONE pointer per network (state_dict order); the modules keep their nn.Parameters (optimiser,
state_dict, copy.deepcopy(mac), freeze_agent_weights, pickling through mp.Queue all keep working, see SURVEY.md 8b
"Ownership").  `ensure_flat` re-points every parameter's `.data` at a slice of one contiguous buffer and repairs the
aliasing whenever something (deepcopy, .to(), load with assign) broke it.
"""
import torch as th


def ensure_flat(module, device=None):
    """Return the flat fp32 buffer that backs all parameters of `module` in registration order."""
    params = getattr(module, "_mal_params", None)
    if params is not None and params:
        owner, attr = module._mal_first       # O(1) probe of the first registered parameter (no generator walk)
        if owner._parameters.get(attr) is not params[0]:
            params = None                      # parameters were re-created (e.g. load_state_dict(assign=True))
            module._mal_total = None
    if params is None:
        params = list(module.parameters())
        module._mal_params = params           # the module structure of this path is fixed after construction
        module._mal_first = None
        for sub in module.modules():
            if sub._parameters:
                name = next(k for k, v in sub._parameters.items() if v is not None)
                module._mal_first = (sub, name)
                break
    if not params:
        return None
    dev = th.device(device) if device is not None else params[0].device
    flat = getattr(module, "_mal_flat", None)
    total = getattr(module, "_mal_total", None)
    if total is None:
        total = module._mal_total = sum(p.numel() for p in params)
    ok = flat is not None and flat.numel() == total and flat.device == dev and flat.dtype == th.float32
    if ok and getattr(module, "_mal_flat_checked", False):
        # fast path (every kernel call): first and last parameter still alias the buffer
        last = params[-1]
        if params[0].data_ptr() == flat.data_ptr() and last.data_ptr() == flat.data_ptr() + 4 * (total - last.numel()):
            return flat
    if ok:
        off = 0
        base = flat.data_ptr()
        for p in params:
            if p.data_ptr() != base + 4 * off or p.dtype != th.float32 or not p.is_contiguous():
                ok = False
                break
            off += p.numel()
    if not ok:
        flat = th.empty(total, dtype=th.float32, device=dev)
        off = 0
        with th.no_grad():
            for p in params:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                off += n
        module._mal_flat = flat
    module._mal_flat_checked = True
    return flat


def flat_views(flat, params, offset=0):
    """Per-parameter views of `flat` (e.g. a gradient or square_avg buffer) shaped like `params`."""
    out, off = [], offset
    for p in params:
        n = p.numel()
        out.append(flat[off:off + n].view(p.shape))
        off += n
    return out
