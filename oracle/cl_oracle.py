"""Dict-of-tensors restatement of the EWC / MAS regulariser hooks.  TEST INFRASTRUCTURE ONLY.

Follows (paths relative to /root/reference):
  get_penalty_grads        cl_baseline_ewc.py:69-81   pen[k] = e_lambda*2*F[k]*(theta[k]-theta*[k]);
                                                      penalty_avg = mean_k mean|pen[k]|
  Fisher accumulation      cl_baseline_ewc.py:245-255 F[k] += mean(loss) * grad[k]**2
  Fisher finalise/merge    cl_baseline_ewc.py:267-282 F /= total_ds ; main = F (first) else gamma*main + F
  MAS penalty              cl_baseline_mas.py:70-75   sum_n sum(Omega[n]*(p_n-theta*[n])**2)
  MAS accumulation         cl_baseline_mas.py:267-270 Omega[n] += |grad_n|
  MAS finalise             cl_baseline_mas.py:283-288 Omega /= len(dataloader); main = Omega (overwrite)
  parameter helpers        utils.py:273-321

PARITY: the reference holds NO tests for any of this ("parity unpinned" by reference tests).  It is
pinned instead to outputs of the reference's own hook functions, whose source is executed in the
authoring container (oracle/ref_import.load_cl_hooks) -> tests/golden/ref_cl.npz.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def get_penalty_grads(e_lambda: float, fish: Dict[str, torch.Tensor], curr: Dict[str, torch.Tensor],
                      checkpoint: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], float]:
    result = {}
    penalty_avg = 0.0
    n = 0
    for key in curr.keys():
        result[key] = e_lambda * 2 * fish[key] * (curr[key] - checkpoint[key])
        penalty_avg = penalty_avg + torch.mean(torch.abs(result[key]))
        n += 1
    return result, float(penalty_avg) / n


def fisher_accumulate(fish: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], loss: torch.Tensor) -> None:
    w = torch.mean(loss.detach().clone())
    for key in list(grads.keys()):
        fish[key] += w * grads[key] ** 2


def fisher_finalise(fish, main_fish, total_ds: int, e_gamma: float):
    for key in fish:
        fish[key] /= total_ds
    if main_fish is None:
        return fish
    for key in fish:
        if main_fish[key] is None:
            main_fish[key] = fish[key]
        else:
            main_fish[key] *= e_gamma
            main_fish[key] += fish[key]
    return main_fish


def mas_penalty(named_params: Dict[str, torch.Tensor], importance, prev_params) -> torch.Tensor:
    loss = 0
    for n, p in named_params.items():
        loss = loss + torch.sum(importance[n] * (p - prev_params[n]) ** 2)
    return loss


def mas_accumulate(importance, grads) -> None:
    for n, g in grads.items():
        if g is not None:
            importance[n] += g.abs().detach()


def mas_finalise(importance, n_batches: int):
    for k in importance.keys():
        importance[k] /= n_batches
    return importance
