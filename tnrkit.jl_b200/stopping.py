"""Stopping criteria -- mirrors /root/reference/src/utility/stopping.jl."""
from __future__ import annotations


class stopcrit:
    def __and__(self, other):
        return MultipleCrit([self, other])

    def info(self, steps, data):
        return ""


class maxiter(stopcrit):
    def __init__(self, n: int):
        self.n = int(n)

    def __call__(self, steps, data):
        return steps < self.n

    def info(self, steps, data):
        return f"Maximum amount of iterations reached: {steps}"


class convcrit(stopcrit):
    def __init__(self, delta: float, f):
        self.delta = float(delta)
        self.f = f

    def __call__(self, steps, data):
        return self.delta < self.f(steps, data)

    def info(self, steps, data):
        return "Convergence criterion reached: %.3e <= %.3e" % (self.f(steps, data), self.delta)


class MultipleCrit(stopcrit):
    def __init__(self, crits):
        self.crits = list(crits)

    def __and__(self, other):
        return MultipleCrit(self.crits + [other])

    def __call__(self, steps, data):
        # continue only while every criterion says continue
        return not any(not c(steps, data) for c in self.crits)

    def info(self, steps, data):
        for c in self.crits:
            if not c(steps, data):
                return c.info(steps, data)
        return ""


def trivial_convcrit(delta):
    return convcrit(delta, lambda steps, data: data[-1])
