from .ewc import fisher_accumulate, fisher_finalise, get_penalty_grads, get_penalty_grads_async
from .flat import FlatDict, FlatParams, Layout, as_flat, flat_params
from .mas import mas_accumulate, mas_finalise, penalty, penalty_into_grads
from .utils import (freeze_layer, get_grads, get_params, get_params_clone, get_zero_params, load_cl_state,
                    save_cl_state, set_grads)

__all__ = [
    "fisher_accumulate", "fisher_finalise", "get_penalty_grads", "get_penalty_grads_async", "FlatDict",
    "FlatParams", "Layout", "as_flat", "flat_params", "mas_accumulate", "mas_finalise", "penalty",
    "penalty_into_grads", "freeze_layer", "get_grads", "get_params", "get_params_clone", "get_zero_params",
    "set_grads", "save_cl_state", "load_cl_state",
]
