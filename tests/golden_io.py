import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def raw_weights(name):
    z = np.load(os.path.join(GOLD, "keras_raw_%s.npz" % name))
    return {k.replace("__", "/"): z[k] for k in z.files}


def load(name):
    return np.load(os.path.join(GOLD, name))
