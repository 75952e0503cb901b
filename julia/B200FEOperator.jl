# B200FEOperator.jl -- Gridap FEOperator plug-in that routes GeMotion's hot path (the per-Newton-iteration assembly of
# /root/reference/src/Solver.jl:142-157) through libgemotion_b200.so (include/gemotion_b200.h).
#
# STATUS: complete but NOT EXECUTED in the build container (no Julia there).  Written against Gridap 0.18.9, the version
# pinned by the reference's Manifest.toml:637-641.  tools/dump_gridap_golden.jl + tests/test_gridap_golden.py check the
# same table extraction against Gridap's own assembly on any machine with Julia.
#
# Use from src/Solver.jl:157 (the only line that changes; `assembly` is the new keyword of GeMotion.simulate, Solver.jl:48-70):
#     op = assembly === :b200 ? B200Assembly.B200FEOperator(X, Y, model, dΩ; Pr, Ra, n, jac_scaling) : FEOperator(res, jac, X, Y)
# Everything above (NLsolve, line search, FEFunction) and below (LUSolver, post-processing) is unchanged.
module B200Assembly

using Gridap, Gridap.FESpaces, Gridap.Algebra, Gridap.Geometry, Gridap.ReferenceFEs, Gridap.CellData, Gridap.Arrays, Gridap.Fields
using SparseArrays

const LIB = get(ENV, "GEMOTION_B200_LIB", "libgemotion_b200.so")
const GM_JACOBIAN = UInt32(1); const GM_RESIDUAL = UInt32(2); const GM_CHECK_FINITE = UInt32(4)
const GM_QUAD = Int32(0); const GM_TRI = Int32(1)

# mirror of gm_desc (include/gemotion_b200.h) -- field order and types must match exactly (checked by struct_size)
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
  kernel::Int32; xchunk::Int32
  n_own_cells::Int64; ghost_cell_ids::Ptr{Int64}   # unstructured partition over several GPUs (0 / NULL: all cells are own cells)
end

gm_check(rc) = rc == 0 || error("gemotion_b200: " * unsafe_string(ccall((:gm_last_error, LIB), Cstring, ())))

"""
Tables the library consumes, extracted from Gridap objects (SURVEY.md 8b).  Also used by tools/dump_gridap_golden.jl.
All arrays are plain `Vector`s in the layouts of include/gemotion_b200.h:
cell dof ids signed, 1-based, field-local, local order U: all u_x nodes then all u_y nodes; P: 3; T: nn.
"""
function extract_tables(X, model, dΩ)
  U, P, T = X.spaces
  grid = get_grid(model)
  poly = first(get_polytopes(model))
  isquad = is_n_cube(poly)
  nn = isquad ? 9 : 6; nv = isquad ? 4 : 3
  ncells = num_cells(model)
  # --- signed cell dof ids (Table{Int32}: data + ptrs); the vector field is permuted to component-major local order
  tU, tP, tT = get_cell_dof_ids(U), get_cell_dof_ids(P), get_cell_dof_ids(T)
  reffeU = ReferenceFE(poly, lagrangian, VectorValue{2,Float64}, 2)
  dofb = get_dof_basis(reffeU)                                  # LagrangianDofBasis: dof_to_node, dof_to_comp
  perm = zeros(Int, 2nn)                                        # our local dof (comp-1)*nn + node  <-  Gridap local dof
  for (ldof, (node, comp)) in enumerate(zip(dofb.dof_to_node, dofb.dof_to_comp))
    perm[(comp - 1) * nn + node] = ldof
  end
  du = Vector{Int32}(undef, ncells * 2nn); dp = Vector{Int32}(undef, ncells * 3); dt = Vector{Int32}(undef, ncells * nn)
  for c in 1:ncells
    ids = tU.data[tU.ptrs[c]:tU.ptrs[c+1]-1]
    @assert length(ids) == 2nn
    du[(c-1)*2nn+1:c*2nn] = ids[perm]
    dp[(c-1)*3+1:c*3] = tP.data[tP.ptrs[c]:tP.ptrs[c+1]-1]
    dt[(c-1)*nn+1:c*nn] = tT.data[tT.ptrs[c]:tT.ptrs[c+1]-1]
  end
  vu = collect(Float64, get_dirichlet_dof_values(U))            # one scalar per Dirichlet dof (u_noslip = 0)
  vp = collect(Float64, get_dirichlet_dof_values(P))            # zeromean: the one fixed dof, value 0
  vt = collect(Float64, get_dirichlet_dof_values(T))
  isempty(vp) && (vp = [0.0])
  n_free = (Int64(num_free_dofs(U)), Int64(num_free_dofs(P)), Int64(num_free_dofs(T)))
  n_diri = (Int64(length(vu)), Int64(max(num_dirichlet_dofs(P), 1)), Int64(length(vt)))
  # --- cell quadrature of Measure(Ω, 2) (Solver.jl:107-109): identical for every cell
  q1 = first(get_data(dΩ.quad))
  qp = get_coordinates(q1); w = collect(Float64, get_weights(q1)); nq = length(w)
  # --- tabulated reference bases at the quadrature points, row-major [q][shape]
  sT = get_shapefuns(ReferenceFE(poly, lagrangian, Float64, 2))
  sP = get_shapefuns(ReferenceFE(poly, lagrangian, Float64, 1; space=:P))
  sG = get_shapefuns(ReferenceFE(poly, lagrangian, Float64, 1))          # geometry (vertex) basis
  vT = evaluate(sT, qp); gT = evaluate(Broadcasting(∇)(sT), qp)          # [nq, nn] values / VectorValue gradients
  vP = evaluate(sP, qp); gG = evaluate(Broadcasting(∇)(sG), qp)
  phi = Float64[vT[q, i] for q in 1:nq for i in 1:nn]
  dphi = Float64[gT[q, i][k] for q in 1:nq for i in 1:nn for k in 1:2]
  psi = Float64[vP[q, m] for q in 1:nq for m in 1:3]
  dgphi = Float64[gG[q, v][k] for q in 1:nq for v in 1:nv for k in 1:2]
  # --- geometry
  cartesian = model isa CartesianDiscreteModel
  nx = ny = 0; origin = (0.0, 0.0); h = (0.0, 0.0)
  coords = Float64[]; cellv = Int32[]
  nverts = num_nodes(grid)
  if cartesian
    desc = get_cartesian_descriptor(model)
    nx, ny = desc.partition
    origin = (Float64(desc.origin[1]), Float64(desc.origin[2])); h = (Float64(desc.sizes[1]), Float64(desc.sizes[2]))
  else
    xs = get_node_coordinates(grid)
    coords = Float64[x[k] for x in xs for k in 1:2]
    tn = get_cell_node_ids(grid)
    cellv = Int32[tn.data[tn.ptrs[c] + v - 1] for c in 1:ncells for v in 1:nv]    # 1-based: vertex_index_base = 1
  end
  (; isquad, nn, nv, nq, ncells, nverts, du, dp, dt, vu, vp, vt, n_free, n_diri, w, phi, dphi, psi, dgphi,
     cartesian, nx, ny, origin, h, coords, cellv)
end

mutable struct B200FEOperator <: FEOperator
  trial::FESpace
  test::FESpace
  handle::Ptr{Cvoid}
  colptr::Vector{Int64}
  rowval::Vector{Int64}
  n::Int
  function B200FEOperator(X, Y, model, dΩ; Pr, Ra, n, jac_scaling=1.0, reg=1e-3, g=(0.0, 1.0), device=0)
    t = extract_tables(X, model, dΩ)
    op = new(X, Y, C_NULL, Int64[], Int64[], num_free_dofs(X))
    href = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve t begin
      desc = Ref(GmDesc(Int32(sizeof(GmDesc)), t.isquad ? GM_QUAD : GM_TRI,
                        Int64(t.ncells), Int64(t.nverts),
                        Int32(t.cartesian), Int32(t.nx), Int32(t.ny), t.origin, t.h,
                        t.cartesian ? Ptr{Float64}(C_NULL) : pointer(t.coords), t.cartesian ? Ptr{Int32}(C_NULL) : pointer(t.cellv),
                        Int32(1), Int32(0),                                   # vertex ids 1-based; GM_LAYOUT_CSC (SparseMatrixCSC)
                        pointer(t.du), pointer(t.dp), pointer(t.dt), t.n_free, t.n_diri,
                        isempty(t.vu) ? Ptr{Float64}(C_NULL) : pointer(t.vu), pointer(t.vp), isempty(t.vt) ? Ptr{Float64}(C_NULL) : pointer(t.vt),
                        Int32(t.nq), Int32(t.nn), Int32(t.nv),
                        pointer(t.w), pointer(t.phi), pointer(t.dphi), pointer(t.psi), pointer(t.dgphi),
                        Int32(device), Int32(0),                              # GM_PATH_AUTO
                        Int32(0), Int32(1), Int64(0), Int64(t.ncells), Int32(0), Int32(0),   # single GPU: rank 0 of 1
                        Int32(0), Int32(0),                                   # GM_KERNEL_AUTO, chunk chosen by the library
                        Int64(0), Ptr{Int64}(C_NULL)))                        # no ghost cells
      gm_check(ccall((:gm_create, LIB), Cint, (Ref{GmDesc}, Ref{Ptr{Cvoid}}), desc, href))
    end
    op.handle = href[]
    finalizer(o -> (o.handle == C_NULL || ccall((:gm_destroy, LIB), Cvoid, (Ptr{Cvoid},), o.handle); o.handle = C_NULL), op)
    gm_check(ccall((:gm_set_params, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble),
                   op.handle, Pr, Ra, n, reg, g[1], g[2], jac_scaling))
    nnz = Ref{Int64}(0)
    gm_check(ccall((:gm_pattern, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), op.handle, nnz))
    op.colptr = Vector{Int64}(undef, op.n + 1); op.rowval = Vector{Int64}(undef, nnz[])
    gm_check(ccall((:gm_get_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int32), op.handle, op.colptr, op.rowval, 1))
    op
  end
end

FESpaces.get_test(op::B200FEOperator) = op.test
FESpaces.get_trial(op::B200FEOperator) = op.trial

# allocate_jacobian: Gridap's SparseMatrixCSC{Float64,Int64} with the library's pattern (1-based colptr / rowval)
nnz_of(op::B200FEOperator) = length(op.rowval)
Algebra.allocate_jacobian(op::B200FEOperator, uh) = SparseMatrixCSC(op.n, op.n, copy(op.colptr), copy(op.rowval), zeros(nnz_of(op)))
Algebra.allocate_residual(op::B200FEOperator, uh) = zeros(op.n)

function _assemble!(op::B200FEOperator, x::Vector{Float64}, nzval, b, flags)
  GC.@preserve x nzval b gm_check(ccall((:gm_assemble, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt32),
      op.handle, x, nzval === nothing ? C_NULL : pointer(nzval), b === nothing ? C_NULL : pointer(b), flags))
end
_x(uh) = convert(Vector{Float64}, get_free_dof_values(uh))     # [U | P | T], contiguous (Solver.jl:178)
Algebra.residual!(b::AbstractVector, op::B200FEOperator, uh) = (_assemble!(op, _x(uh), nothing, b, GM_RESIDUAL); b)
Algebra.jacobian!(A::SparseMatrixCSC, op::B200FEOperator, uh) = (_assemble!(op, _x(uh), A.nzval, nothing, GM_JACOBIAN); A)
function Algebra.residual_and_jacobian!(b::AbstractVector, A::SparseMatrixCSC, op::B200FEOperator, uh)
  _assemble!(op, _x(uh), A.nzval, b, GM_JACOBIAN | GM_RESIDUAL); (b, A)
end
function Algebra.residual_and_jacobian(op::B200FEOperator, uh)
  b = allocate_residual(op, uh); A = allocate_jacobian(op, uh); residual_and_jacobian!(b, A, op, uh)
end
Algebra.residual(op::B200FEOperator, uh) = residual!(allocate_residual(op, uh), op, uh)
Algebra.jacobian(op::B200FEOperator, uh) = jacobian!(allocate_jacobian(op, uh), op, uh)

"Page-lock the arrays NLsolve reuses across iterations (A.nzval, b) so the copies of gm_assemble run at full PCIe speed."
pin!(v::Vector{Float64}) = gm_check(ccall((:gm_host_register, LIB), Cint, (Ptr{Cvoid}, Int64), pointer(v), Int64(sizeof(v))))
unpin!(v::Vector{Float64}) = gm_check(ccall((:gm_host_unregister, LIB), Cint, (Ptr{Cvoid},), pointer(v)))

end # module
