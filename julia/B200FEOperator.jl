# B200FEOperator.jl -- Gridap FEOperator plug-in that routes GeMotion's hot path through
# libgemotion_b200.so (include/gemotion_b200.h).  NOT EXECUTED in the build container (no Julia
# there); written against Gridap 0.18.9 / the pinned Manifest of lamBOOO/GeMotion.jl.
#
# Use from src/Solver.jl:157 --   op = assembly === :b200 ?
#     B200FEOperator(X, Y, model, dΩ; Pr, Ra, n, jac_scaling) : FEOperator(res, jac, X, Y)
# Everything above (NLsolve, line search, FEFunction) and below (LUSolver) is unchanged.
module B200Assembly

using Gridap, Gridap.FESpaces, Gridap.Algebra, Gridap.Geometry, Gridap.ReferenceFEs, Gridap.CellData, Gridap.Arrays
using SparseArrays

const LIB = get(ENV, "GEMOTION_B200_LIB", "libgemotion_b200.so")
const GM_JACOBIAN = UInt32(1); const GM_RESIDUAL = UInt32(2)

# mirror of gm_desc (include/gemotion_b200.h) -- field order and types must match exactly
struct GmDesc
  struct_size::Int32; cell_type::Int32
  ncells::Int64; nverts::Int64
  cartesian::Int32; nx::Int32; ny::Int32
  origin::NTuple{2,Float64}; h::NTuple{2,Float64}
  coords::Ptr{Float64}; cell_vertices::Ptr{Int32}
  vertex_index_base::Int32; layout::Int32
  cell_dofs_u::Ptr{Int32}; cell_dofs_p::Ptr{Int32}; cell_dofs_t::Ptr{Int32}
  n_free::NTuple{3,Int64}; n_diri::NTuple{3,Int64}
  diri_u::Ptr{Float64}; diri_p::Ptr{Float64}; diri_t::Ptr{Float64}
  nq::Int32; nn::Int32; nv::Int32
  w::Ptr{Float64}; phi::Ptr{Float64}; dphi::Ptr{Float64}; psi::Ptr{Float64}; dgphi::Ptr{Float64}
  device::Int32; path::Int32
  rank::Int32; nranks::Int32
  cell_begin::Int64; ncells_global::Int64
  ghost_lo::Int32; ghost_hi::Int32
end

gm_check(rc) = rc == 0 || error("gemotion_b200: " * unsafe_string(ccall((:gm_last_error, LIB), Cstring, ())))

mutable struct B200FEOperator <: FEOperator
  trial::FESpace
  test::FESpace
  handle::Ptr{Cvoid}
  colptr::Vector{Int64}
  rowval::Vector{Int64}
  n::Int
  function B200FEOperator(X, Y, model, dΩ; Pr, Ra, n, jac_scaling=1.0, reg=1e-3, g=(0.0, 1.0), device=0)
    U, P, T = X.spaces
    # --- tables Gridap already holds (SURVEY 8b) ---
    tab(s) = (t = get_cell_dof_ids(s); Int32.(reduce(vcat, collect(t))))     # [ncells*ndofs], signed, 1-based, field-local
    du, dp, dt = tab(U), tab(P), tab(T)
    vu = collect(Float64, reinterpret(Float64, collect(get_dirichlet_dof_values(U))))
    vp = collect(Float64, get_dirichlet_dof_values(P)); vt = collect(Float64, get_dirichlet_dof_values(T))
    quad = CellQuadrature(get_triangulation(dΩ.quad), 2)
    q1 = first(get_data(quad)); qp = get_coordinates(q1); w = collect(Float64, get_weights(q1))
    reffeT = first(get_reffes(T)); reffeP = first(get_reffes(P))
    phi = collect(Float64, transpose(evaluate(get_shapefuns(reffeT), qp)))                 # [nq*nn]
    dphi = Float64[g_[k] for g_ in permutedims(evaluate(Broadcasting(∇)(get_shapefuns(reffeT)), qp)) for k in 1:2]
    psi = collect(Float64, transpose(evaluate(get_shapefuns(reffeP), qp)))
    grid = get_grid(model); isquad = is_n_cube(first(get_polytopes(model)))
    desc_cart = model isa CartesianDiscreteModel
    # ... geometry: Cartesian origin/h from get_cartesian_descriptor(model), else node coordinates,
    #     get_cell_node_ids (1-based -> vertex_index_base = 1) and the geometry reffe gradients `dgphi`
    # (elided: plain array extraction, see INTEGRATION.md)
    op = new(X, Y, C_NULL, Int64[], Int64[], num_free_dofs(X))
    # GC.@preserve du dp dt vu vp vt w phi dphi psi begin
    #   desc = Ref(GmDesc(...)); h = Ref{Ptr{Cvoid}}()
    #   gm_check(ccall((:gm_create, LIB), Cint, (Ref{GmDesc}, Ref{Ptr{Cvoid}}), desc, h)); op.handle = h[]
    # end
    gm_check(ccall((:gm_set_params, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble),
                   op.handle, Pr, Ra, n, reg, g[1], g[2], jac_scaling))
    nnz = Ref{Int64}(0)
    gm_check(ccall((:gm_pattern, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), op.handle, nnz))
    op.colptr = Vector{Int64}(undef, op.n + 1); op.rowval = Vector{Int64}(undef, nnz[])
    gm_check(ccall((:gm_get_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int32), op.handle, op.colptr, op.rowval, 1))
    finalizer(o -> ccall((:gm_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle), op)
    op
  end
end

FESpaces.get_test(op::B200FEOperator) = op.test
FESpaces.get_trial(op::B200FEOperator) = op.trial

# allocate_jacobian: Gridap's SparseMatrixCSC{Float64,Int64} with the library's (bit-exact) pattern
Algebra.allocate_jacobian(op::B200FEOperator, uh) = SparseMatrixCSC(op.n, op.n, copy(op.colptr), copy(op.rowval), zeros(nnz_of(op)))
nnz_of(op) = length(op.rowval)
Algebra.allocate_residual(op::B200FEOperator, uh) = zeros(op.n)

function _assemble!(op, x::Vector{Float64}, nzval, b, flags)
  GC.@preserve x nzval b gm_check(ccall((:gm_assemble, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt32),
      op.handle, x, nzval === nothing ? C_NULL : pointer(nzval), b === nothing ? C_NULL : pointer(b), flags))
end
Algebra.residual!(b::AbstractVector, op::B200FEOperator, uh) = (_assemble!(op, get_free_dof_values(uh), nothing, b, GM_RESIDUAL); b)
Algebra.jacobian!(A::SparseMatrixCSC, op::B200FEOperator, uh) = (_assemble!(op, get_free_dof_values(uh), A.nzval, nothing, GM_JACOBIAN); A)
function Algebra.residual_and_jacobian!(b, A, op::B200FEOperator, uh)
  _assemble!(op, get_free_dof_values(uh), A.nzval, b, GM_JACOBIAN | GM_RESIDUAL); (b, A)
end
function Algebra.residual_and_jacobian(op::B200FEOperator, uh)
  b = allocate_residual(op, uh); A = allocate_jacobian(op, uh); residual_and_jacobian!(b, A, op, uh)
end

end # module
