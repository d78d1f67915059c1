"""Rotation displacement provider -- host mirror of src/rotation.jl:4-31, :59-71."""
from .advection import AbstractExtDataAdv


class RotationVar(AbstractExtDataAdv):
    def __init__(self, adv):
        if adv.N != 2:
            raise ValueError("rotation needs a 2-D grid")
        self.decfl = None

    def initcoef(self, advd):
        """decfl = sign * dt_cur / step(mesh_cur) * mesh_other.points  (src/rotation.jl:21-31):
        kept as (scale, device-resident points) so nothing is uploaded per stage."""
        st_cur, st_other = advd.getst().perm
        mesh_cur = advd.adv.t_mesh[st_cur - 1]
        sign = -1 if st_cur == 1 else 1
        self._scale = sign * advd.getcur_t() / mesh_cur.step
        self._other = st_other
        self.decfl = self._scale * advd.adv.t_mesh[st_other - 1].points

    def initcoef_reads_data(self, advd):
        return False  # shifts come from the meshes / constants only

    def alpha_table(self, advd):  # getalpha(pv, advd, ind) = (decfl[ind],)  (src/rotation.jl:71)
        strides = [0, 0]
        strides[self._other - 1] = 1
        n = advd.adv.sizeall[self._other - 1]
        return (advd.points_dev(self._other - 1), n), strides, self._scale, True


def getrotationvar(adv):
    return RotationVar(adv)
