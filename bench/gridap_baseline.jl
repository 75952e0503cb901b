# bench/gridap_baseline.jl -- BASELINE.md section 3.1 / SURVEY 8(d) "CPU baseline beside it (i)": Gridap's OWN assembly of the hot path
# (residual_and_jacobian! of the operator built at src/Solver.jl:157), timed on the same Cartesian model / BC set / iterate rule as bench.py.
# Julia is not in this repo's image (`julia: not found`), so the script is shipped unexecuted; on a machine with Julia 1.10 + the pinned
# GeMotion manifest:
#     julia --project=/path/to/GeMotion.jl bench/gridap_baseline.jl [N ...]          (default N = 256 512 1024)
# It prints one JSON line per size in bench.py's vocabulary ("impl": "gridap"; Gridap's assembly loop is serial: "cores": 1), and, when
# the library is on LD_LIBRARY_PATH (GEMOTION_B200_LIB), times the same calls through julia/B200FEOperator.jl beside it.
using Gridap, Gridap.FESpaces, Gridap.Algebra, SparseArrays, Statistics, Printf

sizes = isempty(ARGS) ? [256, 512, 1024] : parse.(Int, ARGS)
Pr, Ra, n, reg, g, jac_scaling = 100.0, 1e5, 1.4, 1e-3, VectorValue(0.0, 1.0), 1.0     # bench.py's workload (config 5 parameters)
V_tags, T_tags, T_vals = [1, 2, 3, 4, 5, 6, 7, 8], [7, 8, 1, 2, 3, 4], [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]   # validation-turan/turan.jl:18-23

splitmix(k::UInt64) = (z = k + 0x9E3779B97F4A7C15; z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9; z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB; z ⊻ (z >> 31))

for N in sizes
  model = CartesianDiscreteModel((0, 1, 0, 1), (N, N))
  # --- src/Solver.jl:80-157 ---
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

  nfree = num_free_dofs(X)
  x = [Float64(splitmix(UInt64(1234 + k)) >> 11) * 2.0^-53 for k in 0:nfree-1]      # SURVEY 8(d): splitmix64(seed + k), seed = 1234
  uh = FEFunction(X, x)
  bvec, A = residual_and_jacobian(op, uh)                                               # allocation + compilation + first assembly
  function timed(f; reps=5)
    f()                                                                                  # warm-up
    ts = [(@elapsed f()) for _ in 1:reps]
    median(ts)
  end
  t_both = timed(() -> residual_and_jacobian!(bvec, A, op, uh))
  t_jac = timed(() -> jacobian!(A, op, uh))
  t_res = timed(() -> residual!(bvec, op, uh))
  ncells = N * N
  @printf("{\"impl\": \"gridap\", \"metric\": \"Jacobian+residual assembly Mcells/s (fp64)\", \"value\": %.6g, \"unit\": \"Mcells/s\", \"cores\": 1, \"mesh\": %d, \"ncells\": %d, \"n_free\": %d, \"nnz\": %d, \"ms_per_step\": %.3f, \"jacobian_only_ms\": %.3f, \"residual_only_ms\": %.3f, \"threads\": %d}\n",
          ncells / t_both / 1e6, N, ncells, nfree, nnz(A), 1e3 * t_both, 1e3 * t_jac, 1e3 * t_res, Threads.nthreads())

  if haskey(ENV, "GEMOTION_B200_LIB")                                                         # the drop-in beside it, same calls, host arrays in / out
    include(joinpath(@__DIR__, "..", "julia", "B200FEOperator.jl"))
    opb = B200Assembly.B200FEOperator(X, Y, model, dΩ; Pr=Pr, Ra=Ra, n=n, jac_scaling=jac_scaling)
    bb, Ab = residual_and_jacobian(opb, uh)
    tb = timed(() -> residual_and_jacobian!(bb, Ab, opb, uh))
    err = maximum(abs.(Ab.nzval .- A.nzval)) / maximum(abs.(A.nzval))
    @printf("{\"impl\": \"b200 plug-in\", \"value\": %.6g, \"unit\": \"Mcells/s\", \"mesh\": %d, \"ms_per_step\": %.3f, \"pattern_equal\": %s, \"max_rel_err_nzval\": %.3g}\n",
            ncells / tb / 1e6, N, 1e3 * tb, string(Ab.colptr == A.colptr && Ab.rowval == A.rowval), err)
  end
end
