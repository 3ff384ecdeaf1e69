"""CPU oracle for the RNNT-joint / transducer-loss / CTC / EWC-MAS hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``indic_cl_asr_b200``) never imports from here and fails loudly when its CUDA
library is missing.

Contents
--------
rnnt_oracle.py   numpy float64 restatement of the transducer loss
                 (reference: NeMo/nemo/collections/asr/parts/numba/rnnt_loss/**)
ctc_oracle.py    numpy float64 restatement of CTC forward-backward
                 (reference: NeMo/nemo/collections/asr/losses/ctc.py -> torch.nn.CTCLoss)
joint_oracle.py  torch-CPU restatement of RNNTJoint maths and its fused sub-batch loop
                 (reference: NeMo/nemo/collections/asr/modules/rnnt.py:1375-1665)
cl_oracle.py     dict-of-tensors restatement of the EWC / MAS hooks
                 (reference: cl_baseline_ewc.py, cl_baseline_mas.py, utils.py)
lattice.c        plain-C restatement of the per-utterance alpha/beta/grad loops of
                 cpu_rnnt.py, used as the timed CPU baseline ("port")
ref_import.py    imports the *unmodified* reference (only where /root/reference exists)
gen_golden.py    runs the reference here and freezes tests/golden/*.npz

Parity pinning: rnnt_oracle and ctc_oracle are pinned to the reference's own
golden vectors (tests/golden/ref_kat.npz, extracted from the reference's test
files) and to outputs of the reference itself run in the authoring container
(tests/golden/ref_*.npz).  joint_oracle and cl_oracle are pinned to outputs of
the reference's own functions executed in the authoring container (the
reference has no tests for them).
"""
