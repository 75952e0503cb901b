# dump_gridap_golden.jl -- run on ANY machine with Julia 1.10 + the pinned GeMotion manifest to close
# "parity unpinned" (DESIGN.md section 2): writes Gridap's own pattern / values / dof tables as raw little-endian
# files that tests/test_gridap_golden.py (to be added with the first dump) compares bit-for-bit / to 1e-12.
#   julia --project=/path/to/GeMotion.jl tools/dump_gridap_golden.jl outdir 50   # config 1 (examples/simple.jl)
using Gridap, Gridap.FESpaces, Gridap.Algebra, SparseArrays, Random
outdir = ARGS[1]; N = parse(Int, ARGS[2]); mkpath(outdir)
model = CartesianDiscreteModel((0, 1, 0, 1), (N, N))
Pr, Ra, n, reg, g, jac_scaling = 0.7, 1e3, 1.0, 1e-3, VectorValue(0.0, 1.0), 1.0
# --- exactly src/Solver.jl:80-157 ---
reffeᵤ = ReferenceFE(lagrangian, VectorValue{2,Float64}, 2); reffe_T = ReferenceFE(lagrangian, Float64, 2)
reffeₚ = ReferenceFE(lagrangian, Float64, 1; space=:P)
V = TestFESpace(model, reffeᵤ, conformity=:H1, dirichlet_tags=[1, 2, 3, 4, 5, 6, 7, 8])
Q = TestFESpace(model, reffeₚ, conformity=:L2, constraint=:zeromean)
Θ = TestFESpace(model, reffe_T, conformity=:H1, dirichlet_tags=[5, 7, 8, 1, 2])
U = TrialFESpace(V, VectorValue(0, 0)); P = TrialFESpace(Q); T = TrialFESpace(Θ, [1.0, 0.0, 0.0, 0.5, 0.5])
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
# --- iterate of SURVEY 8(d): splitmix64(seed + k) ---
function splitmix(k::UInt64)
  z = k + 0x9E3779B97F4A7C15; z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9; z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB; z ⊻ (z >> 31)
end
nfree = num_free_dofs(X); x = [Float64(splitmix(UInt64(1234 + k)) >> 11) * 2.0^-53 for k in 0:nfree-1]
uh = FEFunction(X, x); bvec, A = residual_and_jacobian(op, uh)
dump(name, v) = open(io -> write(io, v), joinpath(outdir, name), "w")
dump("colptr.i64", A.colptr); dump("rowval.i64", A.rowval); dump("nzval.f64", A.nzval); dump("b.f64", bvec); dump("x.f64", x)
for (nm, s) in (("U", U), ("P", P), ("T", T))
  t = get_cell_dof_ids(s); dump("cell_dofs_$nm.i32", Int32.(t.data)); dump("cell_dofs_$(nm)_ptrs.i32", Int32.(t.ptrs))
  dump("diri_$nm.f64", collect(Float64, reinterpret(Float64, collect(get_dirichlet_dof_values(s)))))
end
println("n_free = ", nfree, "  nnz = ", nnz(A))
