"""Rotation displacement provider -- host mirror of src/rotation.jl:4-31, :59-71."""
from . import _lib
from .advection import AbstractExtDataAdv
from .unsplit2d import ABTimeAlg_ip, ABTimeAlg_new, DeviceField, NoTimeAlg


class RotationVar(AbstractExtDataAdv):
    def __init__(self, adv):
        if adv.N != 2:
            raise ValueError("rotation needs a 2-D grid")
        self.decfl = None

    def _initcoef_2d(self, advd):
        """initcoef!(pv::RotationVar{T,2}, ::AdvectionData{T,2,timeopt,ABTimeAlg_ip|ABTimeAlg_new})
        -- src/rotation.jl:36-54: bufcur[i, j] = (-dt/dx1 * x2_j, dt/dx2 * x1_i)"""
        adv = advd.adv
        n1, n2 = adv.sizeall
        dt = advd.getcur_t()
        if advd.bufcur is None:
            advd.bufcur = DeviceField(advd.ctx, n1, n2, 2)
        _lib.check(_lib.lib().slb_fill_dec2d(advd.ctx.h, advd.bufcur.ptr, n1, n2, advd.points_dev(1), -dt / adv.t_mesh[0].step,
                                             advd.points_dev(0), dt / adv.t_mesh[1].step))

    def initcoef(self, advd):
        if advd.adv.timealg in (ABTimeAlg_ip, ABTimeAlg_new):
            return self._initcoef_2d(advd)
        if advd.adv.timealg != NoTimeAlg:
            raise NotImplementedError("the reference defines no rotation initcoef! for this time algorithm (src/rotation.jl:21-42)")
        return self._initcoef_1d(advd)

    def _initcoef_1d(self, advd):
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
