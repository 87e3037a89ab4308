"""Mirror of the checkpoint helper of reference ``utils.py:13-43``."""
from collections import OrderedDict


def f_state_dict_wrapper(state_dict, data_parallel=False):
    """Add (data_parallel=True) or strip the ``module.`` prefix DataParallel / DDP put on state-dict keys, so
    fine-tuned reference checkpoints load into wrapped and bare models alike (main.py:391-395, main_kd.py:116-119)."""
    out = OrderedDict()
    for k, v in state_dict.items():
        has = k.startswith("module")
        if data_parallel:
            out[k if has else "module." + k] = v
        else:
            out[k[7:] if has else k] = v
    return out
