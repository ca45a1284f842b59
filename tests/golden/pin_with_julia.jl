# pin_with_julia.jl — pins this repository's oracle (and with it every bit-exact claim of the CUDA path) against the
# REAL OceanTransportMatrixBuilder.jl.  Julia is not available in the build image, so nothing here has been executed
# there; it is the one command a maintainer with Julia runs:
#
#     python tests/golden/npz_to_raw.py                                   # raw twins of tests/golden/case_*.npz
#     julia --project=/path/to/OceanTransportMatrixBuilder.jl tests/golden/pin_with_julia.jl [tests/golden/raw]
#
# For every case it feeds the fixture's INPUTS (the same bytes the oracle and the GPU tests use) to the unmodified
# reference — makegridmetrics (src/gridcellgeometry.jl:265-311), makeindices (src/matrixbuilding.jl:10-24),
# facefluxesfrommasstransport (src/velocities.jl:118-130), transportmatrix (src/matrixbuilding.jl:128-150) — and diffs
# the outputs against the fixture's frozen OUTPUTS:
#   * wet mask chunks, Lwet, N, colptr and rowval of the five matrices, the six face fluxes, thkcello, Z3D: bit for bit;
#   * haversine fields (edge lengths, distances) and nzval with the reference's own geometry: <= 1e-12 relative
#     (Distances.haversine vs the oracle's restatement; north-star tolerance);
#   * nzval again with the FIXTURE's geometry handed to transportmatrix: bit for bit (same inputs, no FMA either side).
# It also writes what the reference produced next to the fixture (`ref_<name>.*`) so that differences can be inspected.
# Exit code 0 = the oracle is pinned.
using OceanTransportMatrixBuilder
using SparseArrays

# the reference reads `.properties["_FillValue"]` and `x |> Array` of YAXArrays: a minimal stand-in
struct Field{T, N} <: AbstractArray{T, N}
    data::Array{T, N}
    properties::Dict{String, Any}
end
Base.size(f::Field) = size(f.data)
Base.getindex(f::Field, i...) = getindex(f.data, i...)
Base.Array(f::Field) = copy(f.data)

readraw(::Type{T}, path, dims...) where {T} = reshape(reinterpret(T, read(path)), dims...) |> collect
bits(a) = reinterpret(UInt64, collect(vec(Float64.(a))))
relerr(a, b) = maximum(abs.(a .- b) ./ max.(abs.(b), floatmin(Float64)); init = 0.0)

function check(ok::Ref{Bool}, name, cond, detail = "")
    println(cond ? "  ok   " : "  FAIL ", name, isempty(detail) ? "" : "  ($detail)")
    cond || (ok[] = false)
end

function pin_case(dir)
    println(dir)
    ok = Ref(true)
    w = split(read(joinpath(dir, "meta.txt"), String))
    nx, ny, nz = parse.(Int, w[1:3])
    fill, upwind, ρs, use3d = parse(Float64, w[5]), w[6] == "1", parse(Float64, w[7]), w[8] == "1"
    f3(n) = readraw(Float64, joinpath(dir, n * ".f64"), nx, ny, nz)
    f2(n) = readraw(Float64, joinpath(dir, n * ".f64"), nx, ny)
    fv(n) = readraw(Float64, joinpath(dir, n * ".f64"), 4, nx, ny)
    f4(n) = readraw(Float64, joinpath(dir, n * ".f64"), nx, ny, 4)
    props = Dict{String, Any}("_FillValue" => fill)
    areacello, volcello = Field(f2("areacello"), Dict{String, Any}()), Field(f3("volcello"), Dict{String, Any}())
    lev = readraw(Float64, joinpath(dir, "lev.f64"), nz)
    gm = makegridmetrics(; areacello, volcello, lon = f2("lon"), lat = f2("lat"), lev, lon_vertices = fv("lon_vertices"),
                         lat_vertices = fv("lat_vertices"))
    ix = makeindices(gm.v3D)
    # ---- indices
    check(ok, "N", ix.N == length(readraw(Int64, joinpath(dir, "Lwet.i64"), :)))
    check(ok, "Lwet", ix.Lwet == readraw(Int64, joinpath(dir, "Lwet.i64"), :))
    check(ok, "wet3D chunks", ix.wet3D.chunks == readraw(UInt64, joinpath(dir, "wet_chunks.u64"), :))
    # ---- geometry
    check(ok, "thkcello (bits)", isequal(bits(gm.thkcello), bits(f3("thkcello"))))
    check(ok, "Z3D (bits)", isequal(bits(gm.Z3D), bits(f3("Z3D"))))
    dirs = (:south, :east, :north, :west)                       # src/gridcellgeometry.jl:304
    for (name, d) in (("edge", gm.edge_length_2D), ("dedge", gm.distance_to_edge_2D), ("dnbr", gm.distance_to_neighbour_2D))
        want = f4(name)
        got = cat((d[s] for s in dirs)...; dims = 3)
        same_nan = isnan.(got) == isnan.(want)
        e = relerr(got[.!isnan.(want)], want[.!isnan.(want)])
        check(ok, "$name (NaN pattern, <= 1e-12)", same_nan && e <= 1e-12, "max rel err $e")
        write(joinpath(dir, "ref_$name.f64"), got)
    end
    # ---- face fluxes: adds only -> bit for bit
    umo, vmo = Field(f3("umo"), props), Field(f3("vmo"), props)
    ϕ = facefluxesfrommasstransport(; umo, vmo, gridmetrics = gm, indices = ix)
    for k in (:east, :west, :north, :south, :top, :bottom)
        check(ok, "ϕ.$k (bits)", isequal(bits(getfield(ϕ, k)), bits(f3("phi_$k"))))
    end
    # ---- matrices
    ρ = use3d ? f3("rho3d") : ρs
    names = (:T => "T", :Tadv => "Tadv", :TκH => "TkH", :TκVML => "TkVML", :TκVdeep => "TkVdeep")
    function compare(tm, tag, exact)
        for (sym, n) in names
            A = getfield(tm, sym)
            cp, rv, nzv = readraw(Int64, joinpath(dir, n * "_colptr.i64"), :), readraw(Int64, joinpath(dir, n * "_rowval.i64"), :),
                          readraw(Float64, joinpath(dir, n * "_nzval.f64"), :)
            check(ok, "$n colptr [$tag]", A.colptr == cp)
            check(ok, "$n rowval [$tag]", A.rowval == rv)
            if length(A.nzval) == length(nzv)
                exact ? check(ok, "$n nzval bits [$tag]", isequal(bits(A.nzval), bits(nzv))) :
                        check(ok, "$n nzval <= 1e-12 [$tag]", relerr(A.nzval, nzv) <= 1e-12, "max rel err $(relerr(A.nzval, nzv))")
            end
            write(joinpath(dir, "ref_$(n)_colptr.i64"), A.colptr); write(joinpath(dir, "ref_$(n)_rowval.i64"), A.rowval)
            write(joinpath(dir, "ref_$(n)_nzval_$tag.f64"), A.nzval)
        end
    end
    mlotst = f2("mlotst")
    compare(transportmatrix(; ϕ, mlotst, gridmetrics = gm, indices = ix, ρ, upwind), "own-geometry", false)
    # the fixture's geometry and fluxes handed to the reference: identical inputs on both sides
    asdict(a) = Dict(s => a[:, :, q] for (q, s) in enumerate(dirs))
    gm2 = merge(gm, (; thkcello = f3("thkcello"), Z3D = f3("Z3D"), edge_length_2D = asdict(f4("edge")),
                     distance_to_edge_2D = asdict(f4("dedge")), distance_to_neighbour_2D = asdict(f4("dnbr"))))
    ϕ2 = (; east = f3("phi_east"), west = f3("phi_west"), north = f3("phi_north"), south = f3("phi_south"), top = f3("phi_top"),
          bottom = f3("phi_bottom"))
    compare(transportmatrix(; ϕ = ϕ2, mlotst, gridmetrics = gm2, indices = ix, ρ, upwind), "fixture-geometry", true)
    return ok[]
end

root = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "raw")
cases = filter(isdir, readdir(root; join = true))
isempty(cases) && error("no cases under $root: run `python tests/golden/npz_to_raw.py` first")
results = [pin_case(c) for c in cases]
println(all(results) ? "ORACLE PINNED: all $(length(results)) cases agree with OceanTransportMatrixBuilder.jl" :
                       "NOT PINNED: $(count(!, results)) of $(length(results)) cases differ")
exit(all(results) ? 0 : 1)
