"""22-joint (SHREC'17 / DHG) and 46-joint (LMDHG, two hands) skeleton graphs and their partitions.

Follows graph/tools.py:7-22,55-60 (adjacency from directed links, column normalisation, the
'spatial' partition = identity / inward / outward), graph/SHRE_graph.py:4-11 and
graph/LMDHG_graph.py:4-40 (edge lists).  Only numpy; nothing here runs per step.
"""
import numpy as np

_SHREC_IN = ((0, 2), (2, 3), (3, 4), (4, 5), (0, 1), (1, 6), (6, 7), (7, 8), (8, 9), (1, 10), (10, 11),
             (11, 12), (12, 13), (1, 14), (14, 15), (15, 16), (16, 17), (1, 18), (18, 19), (19, 20), (20, 21))
_HAND = ((0, 1), (1, 2), (1, 3), (1, 19), (2, 3), (2, 19), (3, 4), (4, 5), (5, 6), (3, 7), (7, 8), (8, 9),
         (9, 10), (7, 11), (11, 12), (12, 13), (13, 14), (11, 15), (15, 16), (16, 17), (17, 18), (15, 19),
         (19, 20), (20, 21), (21, 22))
_LMDHG_IN = _HAND + tuple((a + 23, b + 23) for a, b in _HAND)


def _adj(links, n):
    m = np.zeros((n, n))
    for src, dst in links:
        m[dst, src] = 1.0
    return m


def _norm_cols(m):
    deg = m.sum(axis=0)
    scale = np.divide(1.0, deg, out=np.zeros_like(deg), where=deg > 0)
    return m * scale[None, :]


def _norm_sym(m):
    deg = m.sum(axis=0)
    s = np.where(deg > 0, np.power(np.where(deg > 0, deg, 1.0), -0.5), 0.0)
    return s[:, None] * m * s[None, :]


class Skeleton:
    """Graph(labeling_mode) with the reference's attributes: A, num_node, self_link, inward, outward, neighbor."""

    _inward = ()
    num_node = 0

    def __init__(self, labeling_mode="uniform"):
        n = self.num_node
        self.self_link = [(i, i) for i in range(n)]
        self.inward = list(self._inward)
        self.outward = [(j, i) for i, j in self.inward]
        self.neighbor = self.inward + self.outward
        self.A = self.get_adjacency_matrix(labeling_mode)

    def get_adjacency_matrix(self, labeling_mode=None):
        if labeling_mode is None:
            return self.A
        n = self.num_node
        eye, nb = _adj(self.self_link, n), _adj(self.neighbor, n)
        if labeling_mode == "uniform":
            return _norm_cols(_adj(self.neighbor + self.self_link, n))
        if labeling_mode == "distance*":
            return eye - _norm_cols(nb)
        if labeling_mode == "distance":
            return np.stack((eye, _norm_cols(nb)))
        if labeling_mode == "spatial":
            return np.stack((eye, _norm_cols(_adj(self.inward, n)), _norm_cols(_adj(self.outward, n))))
        if labeling_mode == "DAD":
            return _norm_sym(_adj(self.neighbor + self.self_link, n))
        if labeling_mode == "DLD":
            return eye - _norm_sym(nb)
        raise ValueError(labeling_mode)


class SHRE(Skeleton):
    _inward = _SHREC_IN
    num_node = 22


class LMDHG(Skeleton):
    _inward = _LMDHG_IN
    num_node = 46


# parent joint of every SHREC joint (bone stream, data_process/Hand_Dataset.py:201-202)
SHREC_PARENT = (0, 0, 0, 2, 3, 4, 1, 6, 7, 8, 1, 10, 11, 12, 1, 14, 15, 16, 1, 18, 19, 20)
