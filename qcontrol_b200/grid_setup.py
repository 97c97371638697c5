"""Coordinate and time grids (reference grid_setup.py:4-49), same class names and return values."""
import numpy


class GridConstructor:
    def __init__(self, conf_task):
        self.L = conf_task.L
        self.np = conf_task.np

    def grid_setup(self):
        """-> (dx, x): symmetric equidistant grid of np points spanning L (dx = 1, x = [0] for np == 1)."""
        n = self.np
        if n < 1:
            raise RuntimeError("The number of collocation points 'np' must be positive integers!")
        if n == 1:
            return numpy.float64(1.0), numpy.array([0.0])
        dx = self.L / (n - 1)
        half = numpy.float64(n - 1) * dx / 2.0
        return dx, numpy.array([numpy.float64(i) * dx - half for i in range(n)])


class ForwardTimeGridConstructor:
    def __init__(self, conf_task):
        self.T = conf_task.T
        self.nt = conf_task.fitter.propagation.nt

    def grid_setup(self):
        """-> (dt, [dt * l for l in 0..nt])"""
        dt = self.T / self.nt
        return dt, [dt * l for l in range(self.nt + 1)]
