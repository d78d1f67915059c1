# SemiLagrangianB200.jl -- thin `ccall` veneer over libslb200.so (include/slb200.h).
#
# The Julia host keeps the reference's own objects (UniformMesh, Lagrange / BSplineLU /
# BSplineFFT / Hermite, Advection, the splitting tables: src/SemiLagrangian.jl:44-64) and
# swaps only the mutable run state: `B200AdvectionData` replaces `AdvectionData`
# (src/advection.jl:229-313) and `advection!(::B200AdvectionData)` replaces
# `advection!` (src/advection.jl:594-704).  No CUDA.jl, no kernel DSL: every numerical step
# is one C call.  This file is logic-free on purpose -- Julia is not available in the build
# environment, so every behaviour that needs testing lives in the library or in the Python
# mirror (semilagrangian.jl_b200/slb200), which binds the very same entry points.
module SemiLagrangianB200

using SemiLagrangian
import SemiLagrangian: Advection, AbstractInterpolation, AbstractExtDataAdv, Lagrange, BSplineLU, BSplineFFT,
    Hermite, get_order, getst, getcur_t, getinterp, sizeall, step, points, vec_k_fft

export B200Context, B200AdvectionData, B200PoissonVar, B200RotationVar, B200TranslationVar,
    advection!, getdata, compute_ee, compute_ke, getenergyall

const LIB = get(ENV, "SLB200_LIB", joinpath(@__DIR__, "..", "lib", "libslb200.so"))

struct SlbError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:slb_last_error, LIB), Cstring, ()))
    rc == -1 && throw(ArgumentError(msg))          # SLB_E_ARG == ArgumentError / DomainError of the reference
    throw(SlbError(rc, msg))
end

# ---- context -------------------------------------------------------------------------------
mutable struct B200Context
    h::Ptr{Cvoid}
    function B200Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:slb_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, r))
        c = new(r[])
        # finalizers run in no particular order: objects that live on this context destroy their device parts only
        # while h is still set (see B200AdvectionData), and the context clears h when it goes
        finalizer(x -> (ccall((:slb_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), c)
        return c
    end
end

devalloc(ctx, nbytes) = (r = Ref{Ptr{Cvoid}}(C_NULL);
    check(ccall((:slb_malloc, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx.h, nbytes, r)); r[])
function todevice(ctx, v::Array{Float64})
    p = devalloc(ctx, sizeof(v))
    check(ccall((:slb_memcpy_h2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx.h, p, v, sizeof(v)))
    check(ccall((:slb_sync, LIB), Cint, (Ptr{Cvoid},), ctx.h))
    return p
end

# ---- interpolation objects: the reference's exact-rational tables go across as Float64 -------
kindof(::Lagrange) = 0
kindof(::BSplineLU) = 1
kindof(::BSplineFFT) = 2
kindof(::Hermite) = 3

function interp_handle(ctx::B200Context, interp::AbstractInterpolation{Float64}, n::Integer)
    order = get_order(interp)
    nc = maximum(length(p.coeffs) for p in interp.tabfct)
    coef = zeros(Float64, nc, order + 1)                      # column j = tabfct[j], i.e. row-major rows in C
    for (j, p) in enumerate(interp.tabfct)
        coef[1:length(p.coeffs), j] .= p.coeffs
    end
    nodes = kindof(interp) in (1, 2) ? Float64.(SemiLagrangian.getbspline(order, 0).(1:order)) : Float64[]
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:slb_interp_create, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Float64}, Cint, Ptr{Float64}, Ref{Ptr{Cvoid}}),
        ctx.h, kindof(interp), order, n, coef, nc, isempty(nodes) ? C_NULL : pointer(nodes), r))
    return r[]
end

# ---- run state -------------------------------------------------------------------------------
# B200AdvectionData(adv, data, parext): same signature as AdvectionData (src/advection.jl:244-250);
# `data` is copied to the device, the caller keeps its array (the reference copies as well).
mutable struct B200AdvectionData{T,N}
    adv::Advection{T,N}
    ctx::B200Context
    state_gen::Int
    time_cur::T
    grid::Ptr{Cvoid}
    interps::Vector{Ptr{Cvoid}}
    points_dev::Vector{Ptr{Cvoid}}
    parext::Any
    function B200AdvectionData(adv::Advection{T,N}, data::Array{T,N}, parext; ctx = B200Context(),
        time_init::T = zero(T)) where {T,N}
        size(data) == sizeall(adv) || throw(ArgumentError("size(data)=$(size(data)) it must be $(sizeall(adv))"))
        g = Ref{Ptr{Cvoid}}(C_NULL)
        ext = Int64[sizeall(adv)...]
        check(ccall((:slb_grid_create, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ref{Ptr{Cvoid}}), ctx.h, N, ext, g))
        check(ccall((:slb_grid_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), g[], data))
        interps = [interp_handle(ctx, adv.t_interp[d], ext[d]) for d = 1:N]
        pts = [todevice(ctx, collect(points(adv.t_mesh[d]))) for d = 1:N]
        self = new{T,N}(adv, ctx, 1, time_init, g[], interps, pts, parext)
        finalizer(self) do x          # grid, interpolation handles and mesh nodes go together, before their context
            x.ctx.h == C_NULL && return
            ccall((:slb_grid_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.grid)
            foreach(h -> ccall((:slb_interp_destroy, LIB), Cvoid, (Ptr{Cvoid},), h), x.interps)
            foreach(p -> ccall((:slb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.h, p), x.points_dev)
        end
        return self
    end
end

getst(self::B200AdvectionData) = getst(self.adv, self.state_gen)
getcur_t(self::B200AdvectionData) = getcur_t(self.adv, self.state_gen)

function getdata(self::B200AdvectionData{T,N}) where {T,N}        # src/advection.jl:317
    flush!(self)
    out = Array{T,N}(undef, sizeall(self.adv))
    check(ccall((:slb_grid_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), self.grid, out))
    return out
end

function nextstate!(self::B200AdvectionData)                     # src/advection.jl:358-367
    if self.state_gen < self.adv.nbstates
        self.state_gen += 1
        return true
    end
    self.state_gen = 1
    self.time_cur += self.adv.dt_base
    return false
end

# A displacement provider returns (table::Ptr{Cvoid} | Vector{Float64}, len, strides::Vector{Int64}, scale):
#   alpha(line) = scale * table[1 + sum_d (idx_d - 1) * strides[d]]
# -- the device-side replacement of getalpha(parext, advd, indext) (src/advection.jl:221-224).
const PENDING = IdDict{Any,Any}()   # stage held back for pair fusion, per B200AdvectionData

# slb_sweep flags (include/slb200.h): 1 = SLB_SWEEP_EXACT (the reference's operation order, bitwise),
# 2 = SLB_SWEEP_INSIDE_EDGE (Lagrange(order; edge = InsideEdge) at the kernel seam, src/interpolation.jl:250-286)
sweep_now(self, s) = check(ccall((:slb_sweep, LIB), Cint,
    (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Cdouble, Cint, Cint),
    self.grid, s.dim, self.interps[s.dim+1], s.tab, s.len, s.strides, s.scale, s.ondev, 0))

# run a held-back stage (getdata, compute_ke and every charge density call this first)
function flush!(self::B200AdvectionData)
    s = pop!(PENDING, self, nothing)
    s === nothing || sweep_now(self, s)
end

# Does initcoef! of the CURRENT state read the data?  Default yes; providers overload it
# (Poisson: only the velocity state that recomputes rho, src/poisson.jl:171-176).
initcoef_reads_data(parext, self) = true

# ---- unsplit 2-D states with per-point shifts (src/advection.jl:607-619) -------------------------
# The provider's initcoef! fills BUFCUR[self]: a device buffer [n1, n2, 2] holding the OpTuple{2}
# displacement field as two planes (slb_fill_dec2d for StdPoisson2d / rotation, src/poisson.jl:229-247,
# src/rotation.jl:36-54; slb_lincomb or an upload for user providers such as test/test_swirling.jl:153-167).
# The Adams-Bashforth time algorithms (src/advection.jl:404-580) sequence the same four calls
# (slb_interp2d_points, slb_lincomb, slb_memcpy_d2d, slb_fill_dec2d) exactly as slb200/unsplit2d.py does.
const BUFCUR = IdDict{Any,Ptr{Cvoid}}()

function advection_points!(self::B200AdvectionData{T,2}) where {T}
    (length(self.adv.states) == 1 && getst(self).perm == [1, 2]) ||
        throw(ArgumentError("B200 path: per-point shifts need one unsplit 2-D state ([1, 2], 2, 1, false)"))
    initcoef!(self.parext, self)
    n1, n2 = sizeall(self.adv)
    front = ccall((:slb_grid_front, LIB), Ptr{Cvoid}, (Ptr{Cvoid},), self.grid)
    back = ccall((:slb_grid_back, LIB), Ptr{Cvoid}, (Ptr{Cvoid},), self.grid)
    bs = any(x -> x isa Union{BSplineLU,BSplineFFT}, self.adv.t_interp)
    work = bs ? devalloc(self.ctx, 8 * n1 * n2) : C_NULL
    check(ccall((:slb_interp2d_points, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
        self.ctx.h, self.interps[1], self.interps[2], n1, n2, 1, front, BUFCUR[self], back, work, 0))
    bs && check(ccall((:slb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), self.ctx.h, work))
    check(ccall((:slb_grid_swap, LIB), Cint, (Ptr{Cvoid},), self.grid))
    return nextstate!(self)
end

function advection!(self::B200AdvectionData{T,N}) where {T,N}     # src/advection.jl:594-704
    st = getst(self)
    st.isconstdec || return advection_points!(self)
    st.ndims == 1 || throw(ArgumentError("B200 Julia veneer: const-shift states with ndims = 1 (ndims = 2: see slb200/advection.py)"))
    haskey(PENDING, self) && initcoef_reads_data(self.parext, self) && flush!(self)
    initcoef!(self.parext, self)
    tab, len, strides, scale, ondev = alphatable(self.parext, self)
    cur = (dim = st.perm[1] - 1, tab = tab, len = len, strides = strides, scale = scale, ondev = ondev)
    prev = pop!(PENDING, self, nothing)
    if prev !== nothing && prev.ondev != cur.ondev   # one table on the host, one on the device: not one call
        sweep_now(self, prev); sweep_now(self, cur)
        return nextstate!(self)
    end
    if prev !== nothing
        # two stages in one pass over HBM; bit-identical to two slb_sweep calls
        rc = ccall((:slb_sweep_pair, LIB), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Cdouble,
             Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int64}, Cdouble, Cint, Cint),
            self.grid, prev.dim, self.interps[prev.dim+1], prev.tab, prev.len, prev.strides, prev.scale,
            cur.dim, self.interps[cur.dim+1], cur.tab, cur.len, cur.strides, cur.scale, cur.ondev, 0)
        if rc == -4                       # SLB_E_UNSUPPORTED: not a fusable combination
            sweep_now(self, prev); sweep_now(self, cur)
        else
            check(rc)
        end
        return nextstate!(self)
    end
    if self.state_gen < self.adv.nbstates   # never hold a stage back across the end of a time step
        nxt = getst(self.adv, self.state_gen + 1)
        self.state_gen += 1
        ok = nxt.ndims == 1 && nxt.isconstdec && nxt.perm[1] != 1 && strides[nxt.perm[1]] == 0 &&
             !initcoef_reads_data(self.parext, self)
        self.state_gen -= 1
        if ok
            PENDING[self] = cur
            return nextstate!(self)
        end
    end
    sweep_now(self, cur)
    return nextstate!(self)
end

# ---- Poisson provider (src/poisson.jl:35-224) --------------------------------------------------
mutable struct B200PoissonVar
    Nsp::Int
    plan::Ptr{Cvoid}
    rho::Ptr{Cvoid}
    E::Vector{Ptr{Cvoid}}
    nsp_tot::Int
    cur::Any
    function B200PoissonVar(adv::Advection{T,N}, ctx::B200Context) where {T,N}
        Nsp = div(N, 2)
        fk = SemiLagrangian._get_fctv_k(adv)                       # src/poisson.jl:7-15, purely imaginary
        imags = [collect(imag.(f)) for f in fk]
        ext = Int64[sizeall(adv)[1:Nsp]...]
        ptrs = [pointer(a) for a in imags]
        r = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve imags check(ccall((:slb_poisson_create, LIB), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Ptr{Float64}}, Ref{Ptr{Cvoid}}), ctx.h, Nsp, ext, ptrs, r))
        ntot = prod(ext)
        return new(Nsp, r[], devalloc(ctx, 8ntot), [devalloc(ctx, 8ntot) for _ = 1:Nsp], ntot, nothing)
    end
end

initcoef_reads_data(pv::B200PoissonVar, self::B200AdvectionData) =
    (st = getst(self); st.perm[1] > pv.Nsp && (pv.Nsp + 1) in st.perm[1:st.ndims])

function initcoef!(pv::B200PoissonVar, self::B200AdvectionData{T,N}) where {T,N}   # src/poisson.jl:164-205
    st = getst(self); adv = self.adv; dt = getcur_t(self); Nsp = pv.Nsp
    strides = zeros(Int64, N)
    if st.perm[1] > Nsp
        if (Nsp + 1) in st.perm[1:st.ndims]
            dv = prod(step, adv.t_mesh[(Nsp+1):N])
            flush!(self)
            check(ccall((:slb_charge_density, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}), self.grid, Nsp, dv, pv.rho))
            check(ccall((:slb_poisson_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), pv.plan, pv.rho, pv.E))
        end
        s = 1
        for i = 1:Nsp
            strides[st.perm[N-Nsp+i]] = s
            s *= sizeall(adv)[i]
        end
        pv.cur = (pv.E[st.perm[1]-Nsp], pv.nsp_tot, strides, dt / step(adv.t_mesh[st.perm[1]]), Cint(1))
    else
        tupleind = st.perm[st.invp[1]+Nsp] - st.ndims
        strides[st.perm[st.ndims+tupleind]] = 1
        src = st.invp[1] + Nsp
        pv.cur = (self.points_dev[src], sizeall(adv)[src], strides, -dt / step(adv.t_mesh[st.invp[1]]), Cint(1))
    end
end
alphatable(pv::B200PoissonVar, _) = pv.cur

function compute_ee(self::B200AdvectionData)                      # src/util_poisson.jl:156-162
    pv = self.parext
    dx = prod(step, self.adv.t_mesh[1:pv.Nsp])
    tot = 0.0
    for e in pv.E
        r = Ref{Cdouble}(0)
        check(ccall((:slb_reduce_sumsq, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Cdouble}), self.ctx.h, e, pv.nsp_tot, r))
        tot += r[]
    end
    return dx * tot
end

# ---- rotation / translation providers (src/rotation.jl:21-31,71; src/translation.jl:20-35) -----
mutable struct B200RotationVar
    cur::Any
    B200RotationVar() = new(nothing)
end
function initcoef!(pv::B200RotationVar, self::B200AdvectionData)
    st_cur, st_other = getst(self).perm
    sign = (st_cur == 1) ? -1 : 1
    strides = zeros(Int64, 2); strides[st_other] = 1
    pv.cur = (self.points_dev[st_other], sizeall(self.adv)[st_other], strides,
        sign * getcur_t(self) / step(self.adv.t_mesh[st_cur]), Cint(1))
end
alphatable(pv::B200RotationVar, _) = pv.cur

mutable struct B200TranslationVar{N}
    values::NTuple{N,Float64}
    cur::Any
    B200TranslationVar(v::NTuple{N,Float64}) where {N} = new{N}(v, nothing)
end
function initcoef!(pv::B200TranslationVar{N}, self::B200AdvectionData) where {N}
    st = getst(self)
    pv.cur = ([pv.values[st.perm[1]] * getcur_t(self)], 1, zeros(Int64, N), 1.0, Cint(0))
end
alphatable(pv::B200TranslationVar, _) = (pointer(pv.cur[1]), pv.cur[2], pv.cur[3], pv.cur[4], pv.cur[5])

end # module
