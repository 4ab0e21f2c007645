"""Inputs of the golden vectors (same recipes as tests/golden/make_golden.py)."""
import numpy as np

G = 60.0 * 22050 / 512


def gv1():
    src = [G / 22] * 20 + [G / 21] * 9 + [G / 43] * 4 + [G / 23] * 2
    nc = [G / 17] * 15 + [G / 18] * 8 + [G / 35] * 3 + [G / 16] * 1
    return np.array(nc), np.array(src)


def gv2():
    r = np.random.default_rng(1234)
    a = 120 + r.normal(0, 2, 35)
    b = 150 + r.normal(0, 3, 27)
    return b, a


def gv3():
    r = np.random.default_rng(99)
    s = 0.5 + r.normal(0, .004, 360)
    n = 0.4 + r.normal(0, .004, 361)
    return n, s


def gv4():
    r = np.random.default_rng(7)
    s = 0.5 + r.normal(0, .004, 7200)
    n = 0.4 + r.normal(0, .004, 7201)
    return n, s


GV5_SHIFT = np.array([4, 4, 3, 4, 5, 4, 4]) / 3.0
GV7_A = np.array([.9, .1, .3, .2, .8, .1, .05, .7, .1, .2, .1, .3])
