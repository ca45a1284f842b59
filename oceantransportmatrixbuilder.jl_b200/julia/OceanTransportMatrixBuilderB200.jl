# OceanTransportMatrixBuilderB200.jl — the Julia side of the drop-in boundary.
#
# A thin shim that keeps the reference's exported API (src/OceanTransportMatrixBuilder.jl:31-36)
# and forwards the hot path to libotmb.so (include/otmb.h) with `ccall`.  No CUDA.jl, no kernel
# code generation, no CPU fallback: if the library or a B200 is missing, the calls throw.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image.  The Python
# shim (api.py) is the same layer, exercised by the tests.  See INTEGRATION.md.
module OceanTransportMatrixBuilderB200

using SparseArrays

export makegridmetrics, makeindices, facefluxesfrommasstransport, facefluxesfromvelocities, velocity2fluxes,
       fluxes2velocity, transportmatrix, lump_and_spray

const LIBOTMB = get(ENV, "LIBOTMB", joinpath(@__DIR__, "..", "libotmb.so"))

# the reference's topology structs (src/gridtopology.jl:1-16), reduced to a tag the library understands
abstract type AbstractGridTopology end
struct BipolarGridTopology <: AbstractGridTopology; nx::Int64; ny::Int64; nz::Int64; end
struct TripolarGridTopology <: AbstractGridTopology; nx::Int64; ny::Int64; nz::Int64; end
struct UnknownGridTopology <: AbstractGridTopology; nx::Int64; ny::Int64; nz::Int64; end
topotag(::BipolarGridTopology) = Cint(0)
topotag(::TripolarGridTopology) = Cint(1)
topotag(::UnknownGridTopology) = Cint(2)

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:otmb_create, LIBOTMB), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device)
        st == 0 || error(unsafe_string(ccall((:otmb_status_string, LIBOTMB), Cstring, (Cint,), st)))
        c = new(r[])
        finalizer(c -> ccall((:otmb_destroy, LIBOTMB), Cint, (Ptr{Cvoid},), c.h), c)
        return c
    end
end
const CTX = Ref{Union{Nothing, Context}}(nothing)
ctx() = (CTX[] === nothing && (CTX[] = Context()); CTX[]::Context)

# status -> the reference's exceptions, same messages (src/matrixbuilding.jl:39,61,90,114,233;
# src/gridtopology.jl:111-116; src/velocities.jl:199-200)
function check(c::Context, st::Cint)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:otmb_last_error, LIBOTMB), Cstring, (Ptr{Cvoid},), c.h))
    st == 7 && throw(AssertionError(msg))
    error(msg)
end

isapprox_lon(a, b) = isapprox((@. mod(a - b + 180, 360) - 180), zeros(size(a)), atol = eps(180.0))
function getgridtopology(lon_vertices, lat_vertices, lev)      # src/gridtopology.jl:33-53, host work
    nx, ny, nz = size(lon_vertices, 2), size(lon_vertices, 3), length(lev)
    NPlon = @view lon_vertices[3:4, :, end]
    NPlat = @view lat_vertices[3:4, :, end]
    all(NPlat .== 90) && return BipolarGridTopology(nx, ny, nz)
    (isapprox_lon(NPlon, rot180(NPlon)) && isapprox(NPlat, rot180(NPlat))) && return TripolarGridTopology(nx, ny, nz)
    @warn "Unknown grid topology detected. Things might not work as expected."
    return UnknownGridTopology(nx, ny, nz)
end
function vertexpermutation(lon_vertices, lat_vertices)          # src/gridcellgeometry.jl:158-178, host work
    pts = collect(zip(lon_vertices[:, 1, 1], lat_vertices[:, 1, 1]))
    pe = Set(zip(lon_vertices[:, 2, 1], lat_vertices[:, 2, 1]))
    pn = Set(zip(lon_vertices[:, 1, 2], lat_vertices[:, 1, 2]))
    ie, in_ = findall(in(pe), pts), findall(in(pn), pts)
    i3 = only(ie ∩ in_); i2 = only(setdiff(ie, i3)); i4 = only(setdiff(in_, i3)); i1 = only(setdiff(1:4, i2, i3, i4))
    return [i1, i2, i3, i4]
end

"makegridmetrics, src/gridcellgeometry.jl:265-311"
function makegridmetrics(; areacello, volcello, lon, lat, lev, lon_vertices, lat_vertices)
    toreplace = Set{Any}((missing, nothing, 0))
    haskey(areacello.properties, "_FillValue") && push!(toreplace, areacello.properties["_FillValue"])
    haskey(volcello.properties, "_FillValue") && push!(toreplace, volcello.properties["_FillValue"])
    replacelist = (x => NaN for x in toreplace)
    v3D = Array{Float64}(replace(volcello |> Array{Union{Missing, Float64}}, replacelist...))
    area2D = Array{Float64}(replace(areacello |> Array{Union{Missing, Float64}}, replacelist...))
    zt = Array{Float64}(lev |> Array); lat = Array{Float64}(lat |> Array); lon = Array{Float64}(lon |> Array)
    lon_vertices = lon_vertices |> Array{Float64}; lat_vertices = lat_vertices |> Array{Float64}
    p = vertexpermutation(lon_vertices, lat_vertices)
    lon_vertices = lon_vertices[p, :, :]; lat_vertices = lat_vertices[p, :, :]
    gridtopology = getgridtopology(lon_vertices, lat_vertices, zt)
    nx, ny, nz = size(v3D)
    c = ctx(); N = Ref{Int64}(0)
    check(c, ccall((:otmb_set_grid, LIBOTMB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Cint), c.h, nx, ny, nz, topotag(gridtopology)))
    check(c, ccall((:otmb_makeindices, LIBOTMB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}), c.h, v3D, N))
    thkcello = similar(v3D); Z3D = similar(v3D)
    edge = Array{Float64}(undef, nx, ny, 4); dedge = similar(edge); dnbr = similar(edge)
    check(c, ccall((:otmb_gridmetrics, LIBOTMB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        c.h, area2D, lon, lat, lon_vertices, lat_vertices, zt, thkcello, Z3D, edge, dedge, dnbr))
    dirs = (:south, :east, :north, :west)                      # src/gridcellgeometry.jl:304
    asdict(a) = Dict(d => a[:, :, q] for (q, d) in enumerate(dirs))
    edge_length_2D, distance_to_edge_2D, distance_to_neighbour_2D = asdict(edge), asdict(dedge), asdict(dnbr)
    return (; area2D, v3D, thkcello, lon_vertices, lat_vertices, lon, lat, Z3D, zt, edge_length_2D, distance_to_edge_2D, distance_to_neighbour_2D, gridtopology)
end

"makeindices(v3D), src/matrixbuilding.jl:10-24"
function makeindices(v3D)
    nxyz = size(v3D); M = length(v3D)
    c = ctx(); N = Ref{Int64}(0)
    # the grid tag does not matter for indices; keep whatever makegridmetrics set, else bipolar
    check(c, ccall((:otmb_makeindices, LIBOTMB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}), c.h, v3D, N))
    wet3D = falses(nxyz...)                                     # BitArray chunks are filled in place
    Lwet = Vector{Int64}(undef, N[]); L3 = Array{Int64}(undef, nxyz...)
    check(c, ccall((:otmb_get_indices, LIBOTMB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Int64}, Ptr{Int64}), c.h, wet3D.chunks, Lwet, L3))
    Lwet3D = Array{Union{Int, Missing}, 3}(missing, nxyz...)    # isbits-Union arrays cannot be filled through a pointer
    Lwet3D[Lwet] .= 1:N[]
    return (; wet3D, L = LinearIndices(nxyz), Lwet, N = N[], Lwet3D, C = CartesianIndices(nxyz))
end

"facefluxesfrommasstransport, src/velocities.jl:118-130"
function facefluxesfrommasstransport(; umo, vmo, gridmetrics, indices)
    FillValue = umo.properties["_FillValue"]
    @assert isequal(FillValue, vmo.properties["_FillValue"])
    return facefluxes(umo |> Array{Float64}, vmo |> Array{Float64}, gridmetrics, indices; FillValue)
end

"facefluxes (+ nofluxboundaries!), src/velocities.jl:154-255; umo / vmo are not modified"
function facefluxes(u::Array{Float64, 3}, v::Array{Float64, 3}, gridmetrics, indices; FillValue)
    c = ctx()
    east = similar(u); west = similar(u); north = similar(u); south = similar(u); top = similar(u); bottom = similar(u)
    check(c, ccall((:otmb_facefluxes, LIBOTMB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        c.h, u, v, Float64(FillValue), east, west, north, south, top, bottom))
    return (; east, west, north, south, top, bottom)
end

struct TMParams
    kH::Float64; kVML::Float64; kVdeep::Float64; rho::Float64
    upwind::Int32; index_base::Int32; path::Int32; build_mask::Int32
end

"transportmatrix, src/matrixbuilding.jl:128-150"
function transportmatrix(; ϕ, mlotst, gridmetrics, indices, ρ, κH = 500.0, κVML = 0.1, κVdeep = 1.0e-5,
        Tadv = nothing, TκH = nothing, TκVML = nothing, TκVdeep = nothing, upwind = true)
    c = ctx(); N = indices.N
    (; area2D, thkcello, zt, edge_length_2D, distance_to_neighbour_2D, Z3D, lon, lat) = gridmetrics
    dirs = (:south, :east, :north, :west)
    edge = cat((edge_length_2D[d] for d in dirs)...; dims = 3); dnbr = cat((distance_to_neighbour_2D[d] for d in dirs)...; dims = 3)
    P = Ptr{Float64}
    check(c, ccall((:otmb_set_gridmetrics, LIBOTMB), Cint, (Ptr{Cvoid}, P, P, P, P, P, P, P, P), c.h, area2D, thkcello, zt, edge, dnbr, Z3D, lon, lat))
    faces = [ϕ.east, ϕ.west, ϕ.north, ϕ.south, ϕ.top, ϕ.bottom]
    GC.@preserve faces begin
        check(c, ccall((:otmb_set_facefluxes, LIBOTMB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Float64}}), c.h, pointer.(faces)))
    end
    ml = Array{Float64}(replace(mlotst |> Array, missing => NaN))
    check(c, ccall((:otmb_set_mlotst, LIBOTMB), Cint, (Ptr{Cvoid}, P), c.h, ml))
    if ρ isa Number
        check(c, ccall((:otmb_set_rho3d, LIBOTMB), Cint, (Ptr{Cvoid}, P), c.h, C_NULL))
    else
        check(c, ccall((:otmb_set_rho3d, LIBOTMB), Cint, (Ptr{Cvoid}, P), c.h, Array{Float64}(ρ)))
    end
    pre = (Tadv, TκH, TκVML, TκVdeep); mask = Int32(32)
    for (m, A) in enumerate(pre)
        if isnothing(A)
            mask |= Int32(1) << m
        else
            check(c, ccall((:otmb_set_operator, LIBOTMB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Int64}, P, Int32),
                c.h, m, nnz(A), A.colptr, A.rowval, A.nzval, 1))
        end
    end
    prm = Ref(TMParams(κH, κVML, κVdeep, ρ isa Number ? Float64(ρ) : 0.0, upwind, 1, 0, mask))
    nnzs = zeros(Int64, 5)
    check(c, ccall((:otmb_transportmatrix_build, LIBOTMB), Cint, (Ptr{Cvoid}, Ref{TMParams}, Ptr{Int64}), c.h, prm, nnzs))
    # one pipelined fetch for every matrix that was built (otmb_transportmatrix_fetch_all: Int32 indices on the PCIe
    # link, widened into these Int64 vectors by host threads while the values are in flight)
    want = [true, isnothing(Tadv), isnothing(TκH), isnothing(TκVML), isnothing(TκVdeep)]
    colptrs = [want[m] ? Vector{Int64}(undef, N + 1) : Int64[] for m in 1:5]
    rowvals = [want[m] ? Vector{Int64}(undef, nnzs[m]) : Int64[] for m in 1:5]
    nzvals = [want[m] ? Vector{Float64}(undef, nnzs[m]) : Float64[] for m in 1:5]
    fmask = Cint(sum(want[m] ? 1 << (m - 1) : 0 for m in 1:5))
    GC.@preserve colptrs rowvals nzvals begin
        pc = [want[m] ? pointer(colptrs[m]) : Ptr{Int64}(C_NULL) for m in 1:5]
        pr = [want[m] ? pointer(rowvals[m]) : Ptr{Int64}(C_NULL) for m in 1:5]
        pv = [want[m] ? pointer(nzvals[m]) : Ptr{Float64}(C_NULL) for m in 1:5]
        check(c, ccall((:otmb_transportmatrix_fetch_all, LIBOTMB), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Float64}}),
                       c.h, fmask, pc, pr, pv))
    end
    csc(m) = SparseMatrixCSC{Float64, Int64}(N, N, colptrs[m], rowvals[m], nzvals[m])   # already sorted, 1-based: no copy
    T = csc(1)
    Tadv = isnothing(Tadv) ? csc(2) : Tadv
    TκH = isnothing(TκH) ? csc(3) : TκH
    TκVML = isnothing(TκVML) ? csc(4) : TκVML
    TκVdeep = isnothing(TκVdeep) ? csc(5) : TκVdeep
    return (; T, Tadv, TκH, TκVML, TκVdeep)
end

# ---- velocities <-> mass fluxes (src/velocities.jl:10-108, 140-151).  The Arakawa-grid detection and the
# C/A/B dispatch of interpolateontodefaultCgrid (src/gridcellgeometry.jl:50-140) are host work on one cell and
# keep their Julia code; only the B-grid stencil and the per-cell conversion run on the GPU.
function bgrid_to_cgrid(u::Array{Float64, 3}, v::Array{Float64, 3}, fill)
    c = ctx(); u2 = similar(u); v2 = similar(v)
    check(c, ccall((:otmb_bgrid_to_cgrid, LIBOTMB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}),
                   c.h, u, v, fill, u2, v2))
    return u2, v2
end
function _velflux(sym, a::Array{Float64, 3}, b::Array{Float64, 3}, ρ)
    c = ctx(); oa = similar(a); ob = similar(b)
    ρ3 = ρ isa Number ? Ptr{Float64}(C_NULL) : pointer(ρ)
    GC.@preserve ρ check(c, ccall((sym, LIBOTMB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}),
                                  c.h, a, b, ρ3, ρ isa Number ? Float64(ρ) : 0.0, oa, ob))
    return oa, ob
end
function velocity2fluxes(u, u_lon, u_lat, v, v_lon, v_lat, gridmetrics, ρ)          # src/velocities.jl:10-39
    u, _, _, v, _, _ = interpolateontodefaultCgrid(u, u_lon, u_lat, v, v_lon, v_lat, gridmetrics)   # host dispatch, B-grid -> bgrid_to_cgrid
    return _velflux(:otmb_velocity2fluxes, Array{Float64}(u), Array{Float64}(v), ρ)
end
fluxes2velocity(ϕᵢ, ϕⱼ, gridmetrics, ρ) = _velflux(:otmb_fluxes2velocity, Array{Float64}(ϕᵢ), Array{Float64}(ϕⱼ), ρ)   # :50-74
function facefluxesfromvelocities(; uo, uo_lon, uo_lat, vo, vo_lon, vo_lat, gridmetrics, indices, ρ)   # :140-151
    FillValue = uo.properties["_FillValue"]
    @assert isequal(FillValue, vo.properties["_FillValue"])
    umo, vmo = velocity2fluxes(uo, uo_lon, uo_lat, vo, vo_lon, vo_lat, gridmetrics, ρ)
    return facefluxes(umo, vmo, gridmetrics, indices; FillValue)
end

# ---- lump_and_spray, src/extratools.jl:38-112.  Default mask: on the device; a custom mask makes the reference's
# sweep greedy and data dependent, so that case keeps the reference's Julia body (not repeated here).
function lump_and_spray(wet3D, vol, T::SparseMatrixCSC{Float64, Int64}, mask = trues(size(wet3D)); di = 2, dj = 2, dk = 1)
    all(mask) || error("lump_and_spray with a custom mask: use the reference implementation")
    c = ctx(); N = length(vol); Nc = Ref{Int64}(0)
    check(c, ccall((:otmb_lump_and_spray_build, LIBOTMB), Cint,
                   (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Int32, Int32, Ref{Int64}),
                   c.h, di, dj, dk, vol, T.colptr, T.rowval, 1, 1, Nc))
    lcp = Vector{Int64}(undef, N + 1); lrv = Vector{Int64}(undef, N); lnz = Vector{Float64}(undef, N)
    scp = Vector{Int64}(undef, Nc[] + 1); srv = Vector{Int64}(undef, N); snz = Vector{Float64}(undef, N)
    vol_c = Vector{Float64}(undef, Nc[])
    check(c, ccall((:otmb_lump_and_spray_fetch, LIBOTMB), Cint,
                   (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                   c.h, lcp, lrv, lnz, scp, srv, snz, vol_c))
    return SparseMatrixCSC(Nc[], N, lcp, lrv, lnz), SparseMatrixCSC(N, Nc[], scp, srv, snz), vol_c
end

end # module
