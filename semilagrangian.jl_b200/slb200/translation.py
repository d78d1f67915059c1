"""Translation displacement provider -- host mirror of src/translation.jl:5-35."""
import numpy as np

from .advection import AbstractExtDataAdv


class TranslationVar(AbstractExtDataAdv):
    def __init__(self, values):
        self.values = tuple(float(v) for v in values)
        self.valok = None

    def initcoef(self, advd):  # src/translation.jl:20-23
        st = advd.getst()
        self.valok = tuple(self.values[st.perm[i] - 1] * advd.getcur_t() for i in range(st.ndims))

    def alpha_table_nd(self, advd):
        return [(np.array([v]), [0] * advd.adv.N, 1.0, False) for v in self.valok]

    def initcoef_reads_data(self, advd):
        return False  # shifts come from the meshes / constants only

    def alpha_table(self, advd):  # getalpha: constant shift (src/translation.jl:33-35)
        return np.array([self.valok[0]]), [0] * advd.adv.N, 1.0, False


def gettranslationvar(v):
    return TranslationVar(v)
