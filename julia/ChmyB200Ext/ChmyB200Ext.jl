# ChmyB200Ext -- Julia glue between Chmy.jl's public API and libchmy_b200.so (include/chmy_b200.h).
#
# STATUS: source only.  There is no Julia toolchain in the build image, so this file has never been executed; the
# same C ABI is exercised by the Python host mirror (chmy.jl_b200/) that the test-suite drives.  It documents, in
# the reference's own language, exactly which methods a maintainer adds to make the B200 path a drop-in
# (INTEGRATION.md walks through it).  Struct layouts below must match include/chmy_b200.h; `__init__` verifies
# them against chmy_struct_size() so that a mismatch fails loudly at load time.
module ChmyB200Ext

using Chmy
using Chmy.Architectures, Chmy.Grids, Chmy.Fields, Chmy.BoundaryConditions, Chmy.KernelLaunch, Chmy.Distributed
import Chmy.Architectures: Arch, get_backend, get_device, activate!, set_device!
import Chmy.BoundaryConditions: bc!
import Chmy.Distributed: exchange_halo!
import KernelAbstractions

const libchmy = get(ENV, "CHMY_B200_LIB", "libchmy_b200.so")

# ------------------------------------------------------------------------------------------------ C structs
const MAX_DIMS, MAX_BATCH_FIELDS, MAX_OP_FIELDS, MAX_SCALARS = 3, 8, 24, 8

struct GridDesc                       # chmy_grid_desc
    ndims::Int32
    _pad::Int32
    n::NTuple{3,Int64}
    origin::NTuple{3,Float64}
    extent::NTuple{3,Float64}
    spacing::NTuple{3,Float64}
    inv_spacing::NTuple{3,Float64}
    connectivity::NTuple{6,Int32}     # [dim][side], row-major
end

struct BatchDesc                      # chmy_batch_desc
    kind::Int32
    nfields::Int32
    fields::NTuple{8,Ptr{Cvoid}}
    bc_kind::NTuple{8,Int32}
    value::NTuple{8,Float64}
    value_field::NTuple{8,Ptr{Cvoid}}   # Field-valued conditions (ABI v2); C_NULL -> value
end

struct Inclusion                      # chmy_inclusion
    active::Int32
    loc::NTuple{3,Int32}
    c0::NTuple{3,Float64}
    r::Float64
    in::Float64
    out::Float64
end

struct LaunchDesc                     # chmy_launch_desc
    op::Int32
    flags::Int32
    grid::GridDesc
    nfields::Int32
    nscalars::Int32
    fields::NTuple{24,Ptr{Cvoid}}
    scalars::NTuple{8,Float64}
    rho_g::Inclusion
    has_bc::Int32
    has_outer_width::Int32
    outer_width::NTuple{3,Int64}
    bc::NTuple{6,BatchDesc}           # [dim][side]
    oper::Int32                       # chmy_operator (op == 8 only; ABI v3)
    oper_dim::Int32                   # 0-based
end

function __init__()
    ccall((:chmy_abi_version, libchmy), Cint, ()) == 3 || error("ChmyB200Ext: libchmy_b200.so ABI version mismatch (need 3)")
    for (i, T) in enumerate((GridDesc, BatchDesc, Inclusion, LaunchDesc))
        want = ccall((:chmy_struct_size, libchmy), Csize_t, (Cint,), i - 1)
        want == sizeof(T) || error("ChmyB200Ext: layout of $T ($(sizeof(T)) B) differs from the library ($want B)")
    end
end

check(rc::Integer) = rc == 0 ? nothing :
                     error(unsafe_string(ccall((:chmy_last_error, libchmy), Cstring, ())))   # reference style: plain error()

# ------------------------------------------------------------------------------------------------ backend + architecture
"""The KernelAbstractions-style tag of this path: `Arch(B200Backend(); device_id=1)`."""
struct B200Backend <: KernelAbstractions.Backend end

mutable struct B200Device              # what `get_device` returns; owns the chmy_ctx (device + streams)
    ctx::Ptr{Cvoid}
    id::Int
end

function get_device(::B200Backend, id::Integer)                       # src/Architectures.jl:46-49 via ChmyCUDAExt.jl:17
    ref = Ref{Ptr{Cvoid}}()
    check(ccall((:chmy_ctx_create, libchmy), Cint, (Cint, Ref{Ptr{Cvoid}}), id, ref))
    dev = B200Device(ref[], id)
    finalizer(d -> ccall((:chmy_ctx_destroy, libchmy), Cint, (Ptr{Cvoid},), d.ctx), dev)
    return dev
end
set_device!(::B200Device) = nothing                                   # every entry point selects its own device
KernelAbstractions.synchronize(arch::SingleDeviceArchitecture{B200Backend}) =
    check(ccall((:chmy_synchronize, libchmy), Cint, (Ptr{Cvoid},), get_device(arch).ctx))
KernelAbstractions.priority!(::B200Backend, ::Symbol) = nothing       # stream priorities are fixed inside the ctx
ctx(arch) = get_device(arch).ctx

# ------------------------------------------------------------------------------------------------ fields
"""Device storage handle standing where the `CuArray` stands in `Field{T,N,L,H,A}` (src/Fields/field.jl:6-11)."""
mutable struct B200Array{T<:Union{Float64,Float32},N} <: AbstractArray{T,N}
    handle::Ptr{Cvoid}
    dims::NTuple{N,Int}               # padded dims (dims .+ 4)
    arch
end
Base.size(a::B200Array) = a.dims
dtype_code(::Type{Float64}) = Cint(0)                                  # chmy_dtype
dtype_code(::Type{Float32}) = Cint(1)

function Fields.Field(arch::SingleDeviceArchitecture{B200Backend}, grid::StructuredGrid{N}, loc, ::Type{T}=eltype(grid);
                      halo=1) where {N,T<:Union{Float64,Float32}}      # field.jl:56-74
    halo == 1 || error("the B200 path implements halo = 1")
    loc  = Fields.expand_loc(Val(N), loc)
    dims = size(grid, loc)
    ref  = Ref{Ptr{Cvoid}}()
    check(ccall((:chmy_field_create_typed, libchmy), Cint,
                (Ptr{Cvoid}, Cint, Ref{NTuple{3,Int64}}, Ref{NTuple{3,Int32}}, Cint, Cint, Ref{Ptr{Cvoid}}),
                ctx(arch), N, pad3(dims, 1), pad3(map(l -> Int32(l isa Vertex), loc), 0), 0, dtype_code(T), ref))
    data = B200Array{T,N}(ref[], dims .+ 4, arch)
    finalizer(a -> ccall((:chmy_field_destroy, libchmy), Cint, (Ptr{Cvoid},), a.handle), data)
    return Field{typeof(loc),1}(data, dims)
end

pad3(t::NTuple{N}, fill) where {N} = ntuple(i -> i <= N ? Int64(t[i]) : Int64(fill), 3)
handle(f::Field) = parent(f).handle

# Array(interior(f)) / set!(f, A) / maximum(abs, interior(f)):  sub-box copies and the fused reduction
function Base.Array(f::Field{T,N,L,H,<:B200Array}; with_halo=false) where {T,N,L,H}     # host buffers hold eltype(f)
    lo = ntuple(_ -> with_halo ? 0 : 1, N); hi = size(f) .+ (with_halo ? 1 : 0)
    out = Array{T,N}(undef, (hi .- lo .+ 1)...)
    check(ccall((:chmy_field_copy_to_host, libchmy), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{NTuple{3,Int64}}, Ref{NTuple{3,Int64}}),
                ctx(parent(f).arch), handle(f), out, pad3(lo, 0), pad3(hi, 0)))
    return out
end
function Fields.set!(f::Field{T,N,L,H,<:B200Array}, A::AbstractArray) where {T,N,L,H}     # field.jl:98
    host = Array{T,N}(A)
    check(ccall((:chmy_field_copy_from_host, libchmy), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{NTuple{3,Int64}}, Ref{NTuple{3,Int64}}),
                ctx(parent(f).arch), handle(f), host, pad3(ntuple(_ -> 1, N), 0), pad3(size(f), 0)))
    return
end
function maxabs(f::Field{T,N,L,H,<:B200Array}) where {T,N,L,H}                                 # drivers: maximum(abs.(interior(f)))
    out = Ref{Float64}()                                                                           # exact widening for Float32 fields
    check(ccall((:chmy_field_maxabs, libchmy), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{NTuple{3,Int64}}, Ref{NTuple{3,Int64}}, Ref{Float64}),
                ctx(parent(f).arch), handle(f), pad3(ntuple(_ -> 1, N), 0), pad3(size(f), 0), out))
    return T(out[])
end

# several fields, ONE device round trip (the residual check, stokes_3d_inc_ve_T.jl:171-175): maxabs(∇V, r_V.x, r_V.y, r_V.z)
function maxabs(f1::Field{T,N,L,H,<:B200Array}, fs::Field...) where {T,N,L,H}
    all = (f1, fs...)
    n   = length(all)
    hs  = Ptr{Cvoid}[handle(f) for f in all]
    lo  = Int64[x for f in all for x in pad3(ntuple(_ -> 1, ndims(f)), 0)]
    hi  = Int64[x for f in all for x in pad3(size(f), 0)]
    out = zeros(Float64, n)
    check(ccall((:chmy_field_maxabs_many, libchmy), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                ctx(parent(f1).arch), n, hs, lo, hi, out))
    return ntuple(i -> eltype(all[i])(out[i]), n)
end

# ------------------------------------------------------------------------------------------------ descriptors
function GridDesc(grid::UniformGrid{N}) where {N}
    conn = ntuple(i -> Int32(connectivity(grid, Dim(cld(i, 2)), Side(2 - i % 2)) isa Connected), 2N)
    GridDesc(N, 0, pad3(size(grid, Center()), 0), pad3f(origin(grid, Vertex())), pad3f(extent(grid, Vertex())),
             pad3f(spacing(grid)), pad3f(inv_spacing(grid)), ntuple(i -> i <= 2N ? conn[i] : Int32(0), 6))
end
pad3f(t::NTuple{N}) where {N} = ntuple(i -> i <= N ? Float64(t[i]) : 0.0, 3)

const NO_FIELDS = ntuple(_ -> C_NULL, 8)
const EMPTY_BATCH = BatchDesc(0, 0, NO_FIELDS, ntuple(_ -> Int32(0), 8), ntuple(_ -> 0.0, 8), NO_FIELDS)
BatchDesc(::EmptyBatch, args...) = EMPTY_BATCH
BatchDesc(b::ExchangeBatch, args...) = BatchDesc(2, length(b.fields), ntuple(i -> i <= length(b.fields) ? handle(b.fields[i]) : C_NULL, 8),
                                                 ntuple(_ -> Int32(0), 8), ntuple(_ -> 0.0, 8), NO_FIELDS)

# value(bc, grid, loc, dim, I...) (first_order_boundary_condition.jl:34-40, boundary_function.jl:42-44):
#   nothing -> 0, Number -> value[], lower-dimensional Field -> value_field[] (read at remove_dim(dim, I) by the BC
#   kernel), BoundaryFunction -> evaluated HERE (a closure cannot cross the C ABI) over the face range 0..n_t+2 with the
#   (loc, index) the rule passes (:42-84), uploaded into an (N-1)-dimensional Field that is cached on the arch.
const BF_CACHE = IdDict{Any,Any}()
function boundary_value_field(arch, grid::StructuredGrid{N}, f::Field, c::FirstOrderBC, dim::Dim{D}, side::Side{S}) where {N,D,S}
    get!(BF_CACHE, (arch, grid, c.value, location(f, dim), typeof(c), D, S)) do
        loc_f = location(f, dim)
        d     = size(f, D)
        loc, idx = c isa Dirichlet && loc_f isa Vertex ? (Vertex(), S == 1 ? 1 : d) :
                   c isa Dirichlet                     ? (Center(), S == 1 ? 0 : d + 1) :
                                                         (flip(loc_f), S == 1 ? 0 : d + 1)
        tgrid = UniformGrid(arch; origin=remove_dim(dim, origin(grid, Vertex())), extent=remove_dim(dim, extent(grid, Vertex())),
                            dims=remove_dim(dim, size(grid, Center())))
        vf   = Field(arch, tgrid, Vertex(), eltype(f))
        ext  = remove_dim(dim, size(grid, Vertex()) .+ 2)                  # face points 0..n_t+2 (batch.jl:180-181)
        vals = [c.value(grid, loc, dim, insert_dim(dim, Tuple(J) .- 1, idx)...) for J in CartesianIndices(ext)]
        check(ccall((:chmy_field_copy_from_host, libchmy), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{NTuple{3,Int64}}, Ref{NTuple{3,Int64}}),
                    ctx(arch), handle(vf), eltype(f).(vals), pad3(ntuple(_ -> 0, N - 1), 0), pad3(ext .- 1, 0)))
        vf
    end
end

function BatchDesc(b::FieldBatch, arch, grid, dim, side)               # batch.jl:44-54
    n = length(b.fields)
    kind(c)   = Int32(c isa FirstOrderBC{<:Any,BoundaryConditions.NeumannKind})
    value(c)  = c.value isa Number ? Float64(c.value) : 0.0
    vfield(f, c) = c.value isa AbstractField      ? handle(c.value) :
                   c.value isa BoundaryFunction   ? handle(boundary_value_field(arch, grid, f, c, dim, side)) : C_NULL
    BatchDesc(1, n, ntuple(i -> i <= n ? handle(b.fields[i]) : C_NULL, 8), ntuple(i -> i <= n ? kind(b.conditions[i]) : Int32(0), 8),
              ntuple(i -> i <= n ? value(b.conditions[i]) : 0.0, 8),
              ntuple(i -> i <= n ? vfield(b.fields[i], b.conditions[i]) : C_NULL, 8))
end
batchset(arch, grid, bc::NTuple{N}) where {N} =
    ntuple(i -> i <= 2N ? BatchDesc(bc[cld(i, 2)][2 - i % 2], arch, grid, Dim(cld(i, 2)), Side(2 - i % 2)) : EMPTY_BATCH, 6)

# op registry: the @kernel functions of the example solvers, identified by name; `flatten` orders their arguments as
# include/chmy_b200.h documents for each chmy_op.
const OPS = Dict(:compute_q! => 1, :update_C! => 2, :update_old! => 3, :update_stress! => 4, :update_velocity! => 5,
                 :update_thermal_flux! => 6, :update_thermal! => 7)
flat(x::Field) = (handle(x),)
flat(x::NamedTuple) = mapreduce(flat, (a, b) -> (a..., b...), values(x))     # VectorField / TensorField in declaration order
flat(::FunctionField) = (C_NULL,)
function flatten(args)
    fields  = mapreduce(a -> a isa Union{Field,NamedTuple,FunctionField} ? flat(a) : (), (a, b) -> (a..., b...), args)
    scalars = Tuple(Float64(a) for a in args if a isa Number)
    incl    = something(findfirst(a -> a isa FunctionField, args), 0)
    return fields, scalars, incl == 0 ? Inclusion(0, (0, 0, 0), (0.0, 0.0, 0.0), 0.0, 0.0, 0.0) : Inclusion(args[incl])
end
function Inclusion(f::FunctionField{T,N}) where {T,N}                  # FunctionField(init_incl, grid, loc; parameters=(x0, y0[, z0], r, in, out))
    p = f.parameters
    Inclusion(1, pad3(map(l -> Int32(l isa Vertex), location(f)), 0), pad3f(Tuple(p)[1:N]), p.r, p.in, p.out)
end

# GridOperators as kernels (include/chmy_b200.h: chmy_operator).  User `@kernel`s that call ∂x/lerp/divg/... at an index
# cannot cross the C ABI; the B200 backend offers each operator as a callable kernel object instead:
#     launch(arch, grid, OperatorKernel(:partial, Dim(1)) => (dst, f, grid))        # dst[I] = ∂x(f, grid, I)
#     launch(arch, grid, OperatorKernel(:divg) => (C, V, grid))                      # C[I]   = divg(V, grid, I)
const OPERATORS = Dict(:left => 1, :right => 2, :δ => 3, :∂ => 4, :partial => 4, :∂² => 5, :∂k∂ => 6, :lerp => 7, :hlerp => 8,
                       :divg => 9, :lapl => 10, :divg_grad => 11, :vmag => 12, :grad => 13, :kgrad => 14)
struct OperatorKernel
    oper::Int32
    dim::Int32
end
OperatorKernel(name::Symbol, ::Dim{D}=Dim(1)) where {D} = OperatorKernel(OPERATORS[name], D - 1)
Base.nameof(::OperatorKernel) = :operator!
OPS[:operator!] = 8
operator_of(k::OperatorKernel) = (k.oper, k.dim)
operator_of(_) = (Int32(0), Int32(0))

# Page-locked host arrays for set!(f, A) / Array(interior(f)) at full host-link rate (chmy_host_alloc).
function pinned_array(arch, dims::NTuple{N,Int}, ::Type{T}=Float64) where {N,T<:Union{Float64,Float32}}
    ref = Ref{Ptr{Cvoid}}()
    check(ccall((:chmy_host_alloc, libchmy), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx(arch), prod(dims) * sizeof(T), ref))
    A = unsafe_wrap(Array, Ptr{T}(ref[]), dims)
    finalizer(a -> ccall((:chmy_host_free, libchmy), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), C_NULL, pointer(a)), A)
    return A
end

# ------------------------------------------------------------------------------------------------ the hot entry points
function (launcher::Launcher)(arch::SingleDeviceArchitecture{B200Backend}, grid, kernel_and_args::Pair; bc=nothing)   # KernelLaunch.jl:105-119
    kernel, args = kernel_and_args
    op = get(OPS, nameof(kernel), 0)
    op == 0 && error("the B200 backend runs the named solver kernels $(keys(OPS)) and the operator kernels; got $(nameof(kernel))")
    fields, scalars, incl = flatten(args)
    oper, oper_dim = operator_of(kernel)
    ow = outer_width(launcher)
    desc = LaunchDesc(op, 1 #= BLOCKING: KernelLaunch.jl:117 =#, GridDesc(grid), length(fields), length(scalars),
                      ntuple(i -> i <= length(fields) ? fields[i] : C_NULL, 24), ntuple(i -> i <= length(scalars) ? scalars[i] : 0.0, 8),
                      incl, bc === nothing ? 0 : 1, ow === nothing ? 0 : 1, ow === nothing ? (0, 0, 0) : pad3(ow, 0),
                      bc === nothing ? ntuple(_ -> EMPTY_BATCH, 6) : batchset(arch, grid, bc), oper, oper_dim)
    check(ccall((:chmy_launch, libchmy), Cint, (Ptr{Cvoid}, Ref{LaunchDesc}), ctx(arch), desc))
    return
end

function bc!(arch::SingleDeviceArchitecture{B200Backend}, grid::StructuredGrid, batch::BoundaryConditions.BatchSet)   # batch.jl:20-29
    check(ccall((:chmy_bc, libchmy), Cint, (Ptr{Cvoid}, Ref{GridDesc}, Ref{NTuple{6,BatchDesc}}, Cint),
                ctx(arch), GridDesc(grid), batchset(arch, grid, batch), 1))
    return
end

# Lazily fused launches (include/chmy_b200.h: chmy_set_fusion): the first launch of a pair (`update_stress!`, `compute_q!`,
# `update_thermal_flux!`) is deferred and runs with the following launch of its partner as one sweep; anything else on the
# context flushes it first, so drivers stay exactly as the reference wrote them.  Device pointers cached from
# chmy_field_get_info go stale after a fused launch (the written fields ping-pong between two buffers): B200Array
# re-queries its pointer instead of caching it.
fuse!(arch::SingleDeviceArchitecture{B200Backend}, on::Bool=true) =
    check(ccall((:chmy_set_fusion, libchmy), Cint, (Ptr{Cvoid}, Cint), ctx(arch), on ? 3 : 0))

# sweeps launched / pairs that had to run as two kernels because the device had no room for their shadow buffers
function fused_counts(arch::SingleDeviceArchitecture{B200Backend})
    n, m = Ref{UInt64}(0), Ref{UInt64}(0)
    check(ccall((:chmy_fused_count, libchmy), Cint, (Ptr{Cvoid}, Ref{UInt64}), ctx(arch), n))
    check(ccall((:chmy_fusion_fallback_count, libchmy), Cint, (Ptr{Cvoid}, Ref{UInt64}), ctx(arch), m))
    return (fused=n[], fallbacks=m[])
end

# Launches with boundary batches (include/chmy_b200.h: chmy_set_launch_tuning): overlap the batches / the halo exchange with
# the kernel (default) or run everything on one stream; results are identical.  -1 keeps the batch-folding setting.
overlap!(arch, on::Bool=true) =
    check(ccall((:chmy_set_launch_tuning, libchmy), Cint, (Ptr{Cvoid}, Cint, Cint), ctx(arch), on ? 1 : 0, -1))

# set!(C, grid, init_gauss): the Gaussian initial condition of the diffusion drivers (examples/diffusion_2d_mpi.jl:46)
# evaluated on the device (chmy_field_set_gaussian)
init_gauss(x...) = exp(-sum(abs2, x))
function Fields.set!(f::Field, grid::StructuredGrid, ::typeof(init_gauss))
    check(ccall((:chmy_field_set_gaussian, libchmy), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{GridDesc}), ctx(parent(f).arch), handle(f), GridDesc(grid)))
    return
end

# Transport of exchange_halo! on a distributed architecture (include/chmy_b200.h: chmy_set_exchange_mode): :nccl (default) or
# :peer (opt-in: pack kernels store into the neighbour's HBM over NVLink, sequence flags instead of ncclSend/ncclRecv).
# Every rank must choose the same; results are identical.
exchange_mode!(arch, mode::Symbol) =
    check(ccall((:chmy_set_exchange_mode, libchmy), Cint, (Ptr{Cvoid}, Cint), ctx(arch), mode === :peer ? 1 : 0))

# Distributed: Arch(B200Backend(), comm, dims) builds the CartesianTopology with MPI exactly as the reference does
# (topology.jl:26-41) and then hands rank/size/dims plus an MPI-broadcast NCCL id to chmy_topo_create; after that
# exchange_halo! never touches MPI.
function attach_topology!(arch, topo::CartesianTopology)
    id = zeros(UInt8, 128)
    global_rank(topo) == 0 && check(ccall((:chmy_comm_unique_id, libchmy), Cint, (Ptr{UInt8},), id))
    Distributed.MPI.Bcast!(id, 0, cart_comm(topo))
    d = Int32.(collect(dims(topo)))
    check(ccall((:chmy_topo_create, libchmy), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int32}, Ptr{UInt8}),
                ctx(arch), global_size(topo), global_rank(topo), length(d), d, id))
end

function exchange_halo!(side::Side{S}, dim::Dim{D}, arch::DistributedArchitecture, grid, fields::Vararg{Field,K}) where {S,D,K}   # exchange_halo.jl:13-61
    hs = collect(map(handle, fields))
    check(ccall((:chmy_exchange_halo, libchmy), Cint, (Ptr{Cvoid}, Ref{GridDesc}, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}, Cint),
                ctx(arch), GridDesc(grid), D - 1, S - 1, K, hs, 1))
    return
end
function exchange_halo!(arch::DistributedArchitecture, grid::StructuredGrid, fields::Vararg{Field,K}) where {K}                  # exchange_halo.jl:73-84
    hs = collect(map(handle, fields))
    check(ccall((:chmy_exchange_halo_all, libchmy), Cint, (Ptr{Cvoid}, Ref{GridDesc}, Cint, Ptr{Ptr{Cvoid}}, Cint),
                ctx(arch), GridDesc(grid), K, hs, 1))
    return
end

end # module
