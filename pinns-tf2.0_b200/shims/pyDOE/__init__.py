"""pyDOE.lhs stand-in (classic Latin hypercube: per-dimension stratified uniform + independent permutation), drawing
from numpy's global RNG like pyDOE 0.3.8 does."""
import numpy as np


def lhs(n, samples=None, criterion=None, iterations=None):
    samples = samples or n
    cut = np.linspace(0, 1, samples + 1)
    u = np.random.rand(samples, n)
    a, b = cut[:samples], cut[1:samples + 1]
    rd = u * (b - a)[:, None] + a[:, None]
    H = np.zeros_like(rd)
    for j in range(n):
        H[:, j] = rd[np.random.permutation(range(samples)), j]
    return H
