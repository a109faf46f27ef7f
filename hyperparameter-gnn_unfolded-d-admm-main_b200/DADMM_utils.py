"""Drop-in replacement of the reference ``DADMM_utils.py`` (graph helpers used by the legacy drivers;
off the hot path -- kept importable with the same names, reference DADMM_utils.py:12-104)."""
import os
import random
import sys

import networkx as nx
import numpy as np


class Vars:
    """Problem data holder of the classic D-ADMM scripts: loads ``GaussianData.mat`` (a blob that is
    absent from the reference checkout too, see its .MISSING_LARGE_BLOBS)."""

    def __init__(self, inputs, m_p):
        import scipy.io as sio
        import torch
        self.inputs, self.m_p = inputs, m_p
        here = os.path.realpath(os.path.join(os.getcwd(), os.path.dirname(__file__)))
        mat = os.path.join(here, "GenerateData", "ProblemData", "CompressedSensing", "GaussianData.mat")
        self.A_BPDN = torch.from_numpy(sio.loadmat(mat)["A_BP"])


class CreateGraph:
    def __init__(self, args):
        self.args = args
        self.net1 = nx.erdos_renyi_graph(args.P, args.graph_prob)

    @staticmethod
    def graph2array(net1):
        """Object array of sorted uint8 neighbour arrays, one per node 0..P-1; exits when a node is isolated."""
        out = []
        for node in range(net1.number_of_nodes()):
            nb = sorted(net1.neighbors(node))
            if not nb:
                print("One or more nodes in the graph are not connected\n"
                      "Please increase the probability of the graph and run again")
                sys.exit()
            out.append(np.array(nb, dtype="uint8"))
        return np.array(out, dtype=object)

    @staticmethod
    def proper_coloring_algorithm(network):
        """Greedy proper colouring in random node order; returns the colour classes as uint8 arrays."""
        order = list(network.nodes())
        random.shuffle(order)
        colour = {}
        for node in order:
            taken = {colour[v] for v in network[node] if v in colour}
            c = 0
            while c in taken:
                c += 1
            colour[node] = c
            network.nodes[node]["color"] = c
        classes = {}
        for node in network.nodes():
            classes.setdefault(colour[node], []).append(node)
        return np.array([np.array(v, dtype="uint8") for v in classes.values()], dtype=object)
