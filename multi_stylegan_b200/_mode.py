"""Execution form of the networks' custom autograd Functions.

Default: *first-order* — fused Functions with hand-written backwards (styled.py for the generator, conv.ResBlockFused for
the discriminator's residual blocks).  Their backward passes are not recorded by autograd, so they cannot be
differentiated again.  Inside `higher_order_gradients()` every layer runs through Functions whose backward is composed
of differentiable pieces (conv.py: {forward, dgrad, wgrad} are closed under differentiation; op_static: masked
activation backward, FIR adjoint), which is what the two regularisers of the train step need: R1 differentiates the
discriminator's input gradient (loss.py:311-316), path length the generator's latent gradient
(multi_stylegan_generator.py:193-200).  ModelWrapper and Generator.forward(return_path_length_grads=True) enter the
context themselves."""

import threading

_tls = threading.local()        # the form is a property of the thread that records the forward pass


class higher_order_gradients(object):
    def __enter__(self):
        _tls.depth = getattr(_tls, "depth", 0) + 1
        return self

    def __exit__(self, *exc):
        _tls.depth -= 1
        return False


def higher_order() -> bool:
    return getattr(_tls, "depth", 0) > 0


NO_DOUBLE_BACKWARD = ("multi_stylegan_b200: this fused Function is first-order only; record the forward inside "
                      "multi_stylegan_b200.higher_order_gradients() to differentiate through its backward pass "
                      "(Generator.forward(return_path_length_grads=True) and ModelWrapper's R1 step do so themselves)")
