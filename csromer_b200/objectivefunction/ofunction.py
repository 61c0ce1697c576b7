"""``OFunction`` -- weighted sum of ``Fi`` terms (reference: objectivefunction/ofunction.py:12-62).

The regularisation weight lives in ``F[id].reg`` and is mutated by ``FISTA.run`` (lambda decay).  With a batch,
``reg`` may be a per-pixel array and ``evaluate`` returns one value per pixel.
"""
import numpy as np


class OFunction:

    def __init__(self, F=None):
        self.F = F
        if F is None:
            self.prox_functions = []
            self.nfuncs = None
        else:
            self.values = [0.0] * len(F)
            self.nfuncs = len(F)
            self.prox_functions = list(F)

    def getProxFunctions(self):
        return self.prox_functions

    def getValues(self):
        return self.values

    def getLambda(self, _id=0):
        return self.F[_id].reg

    def setLambda(self, reg=0.0, _id=0):
        self.F[_id].reg = reg

    def evaluate(self, x):
        total = 0.0
        for i, term in enumerate(self.F):
            self.values[i] = term.evaluate(x)
            total = total + term.reg * self.values[i]
        return total

    def calculate_gradient(self, x):
        grad = np.zeros(np.shape(x), dtype=x.dtype)
        for term in self.F:
            grad = grad + term.reg * term.calculate_gradient(x)
        return grad

    def calc_prox(self, x, nu=0, _id=0):
        if len(self.prox_functions) == 1:
            return self.F[_id].calculate_prox(x, nu)
        out = x
        for term in self.F:  # sequential, `nu` ignored -- as in the reference (ofunction.py:58-61)
            out = term.calculate_prox(out)
        return out
