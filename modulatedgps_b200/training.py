"""run_adam — utils/training_utils.py:4-28 with the same signature, printing and return value.

TF's Adam defaults are kept (beta1 0.9, beta2 0.999, epsilon 1e-7).  The optimiser state lives on the device and
updates the unconstrained variables in place (torch.optim.Adam, foreach implementation); the loss and all its
gradients come from libmgp's hand-written forward/backward kernels through `loss.backward()`.
"""
from __future__ import annotations

import torch


def make_adam(model, lr):
    return torch.optim.Adam(list(model.trainable_variables), lr=lr, betas=(0.9, 0.999), eps=1e-7)


def run_adam(model, num_iter, train_iter, lr, compile=True):
    training_loss = model.training_loss_closure(train_iter, compile=compile)
    optimizer = make_adam(model, lr)

    def optimization_step():
        optimizer.zero_grad(set_to_none=True)
        loss = training_loss()
        loss.backward()
        optimizer.step()

    print('{:>5s}'.format("iter") + '{:>24s}'.format("ELBO:"))
    iters = []
    elbos = []
    for i in range(1, num_iter + 1):
        try:
            optimization_step()

            if i % 5 == 0 or i == 0:
                elbo = -float(training_loss().detach())     # a second forward on a NEW minibatch, as the reference
                print('{:>5d}'.format(i) + '{:>24.6f}'.format(elbo))
                iters.append(i)
                elbos.append(elbo)
        except KeyboardInterrupt:
            print("stopping training")
            break

    return iters, elbos
