"""Generate tests/golden/loss_grad.npz: gradients of the reference's own SiSNRWavLoss (imported unmodified) with respect
to the predictions, by torch autograd on CPU, on the inputs of tests/golden/loss.npz.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):  python -m oracle.gen_loss_grad_golden
This is what `batch["loss"].backward()` (src/trainer/trainer.py:46) leaves on s1_pred / s2_pred.
"""
import os

import numpy as np
import torch

from . import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref = ref_import.load()
    z = np.load(os.path.join(GOLDEN, "loss.npz"))
    d = {}
    for i in range(3):
        s1, s2 = torch.from_numpy(z[f"c{i}.s1"]), torch.from_numpy(z[f"c{i}.s2"])
        s1p = torch.from_numpy(z[f"c{i}.s1p"]).clone().requires_grad_(True)
        s2p = torch.from_numpy(z[f"c{i}.s2p"]).clone().requires_grad_(True)
        loss = ref.SiSNRWavLoss()(s1_pred=s1p, s2_pred=s2p, s1=s1, s2=s2)["loss"]
        (3.0 * loss).backward()          # a non-trivial upstream gradient
        d[f"c{i}.g1"], d[f"c{i}.g2"] = s1p.grad.numpy(), s2p.grad.numpy()
        d[f"c{i}.upstream"] = np.array(3.0)
        print(i, float(loss), float(s1p.grad.abs().mean()), float(s2p.grad.abs().mean()))
    np.savez_compressed(os.path.join(GOLDEN, "loss_grad.npz"), **d)


if __name__ == "__main__":
    main()
