# dump_gridap_golden.jl -- run on ANY machine with Julia 1.10 + the pinned GeMotion manifest to close "parity unpinned"
# (DESIGN.md section 2): writes Gridap's OWN assembly (pattern, values, residual) together with every table the C ABI
# consumes, as raw little-endian files + meta.json, for tests/test_gridap_golden.py (GM_GRIDAP_DUMP=<dir of dumps>).
#
#   julia --project=/path/to/GeMotion.jl tools/dump_gridap_golden.jl OUT simple  50  1.0     # config 1 (examples/simple.jl:12-37)
#   julia --project=...                  tools/dump_gridap_golden.jl OUT turan   256 0.6     # config 2 (validation-turan/turan.jl:14-32)
#   julia --project=...                  tools/dump_gridap_golden.jl OUT wave    62  0.6     # study1 (study1.jl:27-43), the CSV golden mesh
#   julia --project=...                  tools/dump_gridap_golden.jl OUT annulus path/to/co-annulus_unstructured_4.msh 1.0   # config 3 (kuehn.jl:13-52)
# Each call writes OUT/<case>_<size>_n<n>/ .  The weak form below is src/Solver.jl:111-152 verbatim (it has to be the
# reference's form to dump the reference's numbers).
using Gridap, Gridap.FESpaces, Gridap.Algebra, Gridap.Geometry, SparseArrays
include(joinpath(@__DIR__, "..", "julia", "B200FEOperator.jl"))   # extract_tables: the plug-in's own table extraction

outroot, case, sizearg, narg = ARGS[1], ARGS[2], ARGS[3], ARGS[4]
n = parse(Float64, narg)
reg, g, jac_scaling = 1e-3, VectorValue(0.0, 1.0), 1.0
if case == "annulus"
  using GridapGmsh
  model = GmshDiscreteModel(sizearg)                                   # kuehn.jl:13-22
  V_tags, T_tags, T_vals = ["all"], ["inner", "outer"], [1.0, 0.0]     # kuehn.jl:47-50
  Pr, Ra = 0.706, 4.7e4
  tag = "annulus_" * splitext(basename(sizearg))[1]
else
  N = parse(Int, sizearg)
  model = CartesianDiscreteModel((0, 1, 0, 1), (N, N))
  V_tags = [1, 2, 3, 4, 5, 6, 7, 8]
  if case == "simple"      # examples/simple.jl:32-36
    T_tags, T_vals, Pr, Ra = [5, 7, 8, 1, 2], [1.0, 0.0, 0.0, 0.5, 0.5], 0.7, 1e3
  elseif case == "turan"   # validation-turan/turan.jl:18-23
    T_tags, T_vals, Pr, Ra = [7, 8, 1, 2, 3, 4], [0.0, 1.0, 0.0, 1.0, 0.0, 1.0], 100.0, 1e5
  elseif case == "wave"    # study1/study1.jl:38-43
    T_tags, T_vals, Pr, Ra = [5, 7, 8, 1, 2], [x -> sin(pi * x[1]), 0.0, 0.0, 0.0, 0.0], 100.0, 1e4
  else
    error("unknown case $case")
  end
  tag = "$(case)_$(N)"
end
outdir = joinpath(outroot, "$(tag)_n$(n)"); mkpath(outdir)

# --- exactly src/Solver.jl:80-157 ---
reffeᵤ = ReferenceFE(lagrangian, VectorValue{2,Float64}, 2); reffe_T = ReferenceFE(lagrangian, Float64, 2)
reffeₚ = ReferenceFE(lagrangian, Float64, 1; space=:P)
V = TestFESpace(model, reffeᵤ, conformity=:H1, dirichlet_tags=V_tags)
Q = TestFESpace(model, reffeₚ, conformity=:L2, constraint=:zeromean)
Θ = TestFESpace(model, reffe_T, conformity=:H1, dirichlet_tags=T_tags)
U = TrialFESpace(V, VectorValue(0, 0)); P = TrialFESpace(Q); T = TrialFESpace(Θ, T_vals)
Y = MultiFieldFESpace([V, Q, Θ]); X = MultiFieldFESpace([U, P, T])
dΩ = Measure(Triangulation(model), 2)
D(∇u) = 1 / 2 * (∇u + ∇u'); gamma_dot(∇u) = reg + 2 * (D(∇u) ⊙ D(∇u))
μ(∇u, n) = (x -> x^((n - 1) / 2)) ∘ gamma_dot(∇u)
dμ(∇u, d∇u, n) = ((n - 1) / 2) * ((x -> x^((n - 3) / 2)) ∘ gamma_dot(∇u)) * 4 * (D(∇u) ⊙ D(d∇u))
conv(u, ∇u) = (∇u') ⋅ u; dconv(du, ∇du, u, ∇u) = conv(u, ∇du) + conv(du, ∇u)
a((u, p, T), (v, q, θ)) = ∫(0 * Pr * ∇(v) ⊙ ∇(u) - (∇ ⋅ v) * p + q * (∇ ⋅ u) + ∇(T) ⋅ ∇(θ) - Ra * Pr * (T) * (g ⋅ v))dΩ
b(u, v) = ∫(2 * Pr * μ(∇(u), n) * D(∇(v)) ⊙ D(∇(u)))dΩ
db(u, du, v) = ∫(2 * Pr * (dμ(∇(u), ∇(du), n) * D(∇(u)) + μ(∇(u), n) * D(∇(du))) ⊙ D(∇(v)))dΩ
c(u, v) = ∫(v ⊙ (conv ∘ (u, ∇(u))))dΩ; dc(u, du, v) = ∫(v ⊙ (dconv ∘ (du, ∇(du), u, ∇(u))))dΩ
d(u, T, θ) = ∫((u ⋅ ∇(T)) * θ)dΩ; dd(u, du, T, dT, θ) = ∫(((du ⋅ ∇(T)) * θ + (u ⋅ ∇(dT)) * θ))dΩ
res((u, p, T), (v, q, θ)) = a((u, p, T), (v, q, θ)) + b(u, v) + c(u, v) + d(u, T, θ)
jac((u, p, T), (du, dp, dT), (v, q, θ)) = jac_scaling * (a((du, dp, dT), (v, q, θ)) + db(u, du, v) + dc(u, du, v) + dd(u, du, T, dT, θ))
op = FEOperator(res, jac, X, Y)

# --- iterate of SURVEY 8(d): splitmix64(seed + k), bit-reproducible in any language ---
function splitmix(k::UInt64)
  z = k + 0x9E3779B97F4A7C15; z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9; z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB; z ⊻ (z >> 31)
end
nfree = num_free_dofs(X); x = [Float64(splitmix(UInt64(1234 + k)) >> 11) * 2.0^-53 for k in 0:nfree-1]
uh = FEFunction(X, x); bvec, A = residual_and_jacobian(op, uh)

t = B200Assembly.extract_tables(X, model, dΩ)
put(name, v) = open(io -> write(io, v), joinpath(outdir, name), "w")
put("colptr.i64", Int64.(A.colptr)); put("rowval.i64", Int64.(A.rowval)); put("nzval.f64", A.nzval); put("b.f64", bvec); put("x.f64", x)
put("cell_dofs_U.i32", t.du); put("cell_dofs_P.i32", t.dp); put("cell_dofs_T.i32", t.dt)
put("diri_U.f64", t.vu); put("diri_P.f64", t.vp); put("diri_T.f64", t.vt)
put("w.f64", t.w); put("phi.f64", t.phi); put("dphi.f64", t.dphi); put("psi.f64", t.psi); put("dgphi.f64", t.dgphi)
t.cartesian || (put("coords.f64", t.coords); put("cell_vertices.i32", t.cellv))
open(joinpath(outdir, "meta.json"), "w") do io
  print(io, """{"source": "Gridap", "case": "$tag", "cell_type": "$(t.isquad ? "QUAD" : "TRI")", "cartesian": $(t.cartesian), "nx": $(t.nx), "ny": $(t.ny),
 "origin": [$(t.origin[1]), $(t.origin[2])], "h": [$(t.h[1]), $(t.h[2])], "ncells": $(t.ncells), "nverts": $(t.nverts),
 "n_free": [$(t.n_free[1]), $(t.n_free[2]), $(t.n_free[3])], "n_diri": [$(t.n_diri[1]), $(t.n_diri[2]), $(t.n_diri[3])],
 "nq": $(t.nq), "nn": $(t.nn), "nv": $(t.nv), "index_base": 1, "vertex_index_base": 1, "nnz": $(nnz(A)),
 "params": {"Pr": $Pr, "Ra": $Ra, "n": $n, "reg": $reg, "g": [0.0, 1.0], "jac_scaling": $jac_scaling}}""")
end
println("wrote ", outdir, ": n_free = ", nfree, "  nnz = ", nnz(A))
