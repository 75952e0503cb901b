/*
 * gemotion_b200.h -- C ABI of the B200-native Navier-Stokes-Fourier assembly library
 * (libgemotion_b200.so, built from gemotion.jl_b200/csrc).
 *
 * The reference (lamBOOO/GeMotion.jl) has no FFI of its own: its hot path is entered at
 *   /root/reference/src/Solver.jl:157      op = FEOperator(res, jac, X, Y)
 * and executed by Gridap 0.18.9's FEOperatorFromWeakForm / SparseMatrixAssembler through
 *   allocate_jacobian / jacobian! / residual! / residual_and_jacobian!
 * which NLsolve calls once per Newton iteration / line-search step (Solver.jl:162-169).
 * The entry points below are what a Gridap `FEOperator` plug-in binds with `ccall`
 * (INTEGRATION.md shows the Julia stub).  Each function names the reference interface it
 * replaces.  Plain pointers and sizes only; no C++ / torch types cross this boundary.
 *
 * Conventions
 *  - Return value: 0 on success, negative gm_status otherwise; text via gm_last_error()
 *    (thread-local).  No exception, callback or signal crosses the boundary.
 *  - Ownership: the caller owns every array it passes.  gm_create copies its inputs to the
 *    device and keeps no host pointer.  Outputs are caller-allocated and written in place.
 *  - A handle is one assembly context on one CUDA device; calls on one handle must be
 *    serialised by the caller; distinct handles are independent.
 *  - There is NO CPU fallback: every entry point fails with GM_ERR_CUDA when no sm_100 device
 *    is usable.
 *
 * Data conventions (Gridap's)
 *  - cell dof ids are signed, 1-based and FIELD-LOCAL: id>0 free dof, id<0 Dirichlet dof whose
 *    value is diri_X[-id-1].  Local dof order inside a cell: U: node + nn*comp (all ux, then
 *    all uy), P: 3, T: nn.  Global free vector x = [U | P | T]  (Solver.jl:104-105,178).
 *  - gradient convention  G[k][a] = d_k u_a  (Gridap).
 */
#ifndef GEMOTION_B200_H
#define GEMOTION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_VERSION 110 /* 0.1.1: gm_desc grew (n_own_cells, ghost_cell_ids), GM_PATH_GATHER */

enum gm_status {
  GM_OK = 0,
  GM_ERR_ARG = -1,      /* bad argument / inconsistent tables */
  GM_ERR_CUDA = -2,     /* CUDA runtime error or no usable device */
  GM_ERR_OOM = -3,      /* device or host allocation failed */
  GM_ERR_STATE = -4,    /* call order (e.g. gm_assemble before gm_pattern) */
  GM_ERR_NCCL = -5,     /* NCCL error */
  GM_ERR_NONFINITE = -6 /* non-finite value detected (only with GM_CHECK_FINITE) */
};

enum gm_cell_type { GM_QUAD = 0, GM_TRI = 1 };

/* Layout of the value array handed back.  Gridap's default matrix is
 * SparseMatrixCSC{Float64,Int64}; the pattern is structurally symmetric, so the
 * integer arrays are identical in both layouts and only the value order differs. */
enum gm_layout { GM_LAYOUT_CSC = 0, GM_LAYOUT_CSR = 1 };

/* gm_assemble flags */
#define GM_JACOBIAN 1u     /* jacobian!(A, op, x)   */
#define GM_RESIDUAL 2u     /* residual!(b, op, x)   */
#define GM_CHECK_FINITE 4u /* scan nzval and the residual for NaN/Inf on the device (both entry points; the device one then synchronises) */
#define GM_NO_EXCHANGE 8u  /* multi-GPU: skip the NCCL interface-row sum (caller moves the packed buffers) */

/* Numeric kernels */
enum gm_path {
  GM_PATH_AUTO = 0,    /* structured if the tables verify as a Cartesian strip mesh, else gather */
  GM_PATH_SCATTER = 1, /* colour-ordered atomic-free scatter, any mesh (first generation, kept for A/B) */
  GM_PATH_CART = 2,    /* structured owner-computes row gather, Cartesian quads only */
  GM_PATH_GATHER = 3   /* cell matrices by node pairs + owner-computes list gather, any mesh (Gmsh triangles) */
};

/* Kernel generation of the structured path (A/B measurements, parity tests of the older generations) */
enum gm_kernel {
  GM_KERNEL_AUTO = 0,  /* the fastest one: GM_KERNEL_MMA */
  GM_KERNEL_MMA = 1,   /* fp64 tensor-pipe formulation (k_mma) */
  GM_KERNEL_GATH = 2,  /* register gather (k_gath), two trios of matrix warps */
  GM_KERNEL_GATH1 = 3, /* register gather, one trio */
  GM_KERNEL_SWEEP = 4  /* first generation shared-memory sweep (k_cart) */
};

typedef struct gm_handle gm_handle;

/*
 * Problem description.  Replaces what Gridap holds inside X, Y, dΩ and the assembler
 * (Solver.jl:80-109): get_cell_dof_ids per field, get_dirichlet_dof_values, num_free_dofs,
 * the grid, the cell quadrature and the tabulated reference bases.
 */
typedef struct gm_desc {
  int32_t struct_size; /* = sizeof(gm_desc), for ABI versioning */
  int32_t cell_type;   /* gm_cell_type */
  int64_t ncells;
  int64_t nverts;
  /* geometry: either Cartesian (cells x-fastest, J = diag(h) exactly like Gridap's
   * CartesianDiscreteModel) or vertex coordinates + cell->vertex ids */
  int32_t cartesian;
  int32_t nx, ny;
  double origin[2];
  double h[2];
  const double *coords;         /* [nverts*2], NULL if cartesian */
  const int32_t *cell_vertices; /* [ncells*nv], NULL if cartesian */
  int32_t vertex_index_base;    /* 0 or 1 */
  int32_t layout;               /* gm_layout of nzval */
  /* dof tables */
  const int32_t *cell_dofs_u; /* [ncells*2nn] */
  const int32_t *cell_dofs_p; /* [ncells*3]   */
  const int32_t *cell_dofs_t; /* [ncells*nn]  */
  int64_t n_free[3];          /* U, P, T */
  int64_t n_diri[3];
  const double *diri_u, *diri_p, *diri_t;
  /* reference element + quadrature (Measure(Ω,2), Solver.jl:107-109) */
  int32_t nq, nn, nv; /* quadrature points, scalar P2 nodes (9|6), geometry nodes (4|3) */
  const double *w;     /* [nq]       */
  const double *phi;   /* [nq*nn]    scalar P2 shape values   */
  const double *dphi;  /* [nq*nn*2]  reference gradients      */
  const double *psi;   /* [nq*3]     pressure shape values    */
  const double *dgphi; /* [nq*nv*2]  geometry shape gradients (ignored if cartesian) */
  /* execution */
  int32_t device; /* CUDA device ordinal */
  int32_t path;   /* gm_path */
  /* multi-GPU row-block partition (one process per GPU).  This rank holds the cells
   * [cell_begin, cell_end) of a mesh of ncells_global cells; its tables above describe only
   * those cells but use GLOBAL dof ids.  Single GPU: rank 0 of 1, all cells. */
  int32_t rank, nranks;
  int64_t cell_begin, ncells_global;
  /* Cartesian row blocks: the local table holds `ghost_lo` ghost cell rows below and `ghost_hi`
   * above the owned rows (0 or 1 each; ny counts all local rows).  Ghost cells only complete the
   * symbolic pattern of the interface rows; they are never assembled. */
  int32_t ghost_lo, ghost_hi;
  /* structured path: kernel generation (gm_kernel) and quad columns per x-chunk (0 = chosen from the grid size; tests
   * force 32 / 128 / 256 to put chunk seams on small meshes) */
  int32_t kernel, xchunk;
  /* Unstructured meshes on several GPUs (GM_PATH_GATHER): contiguous global cell ranges per rank.  The first
   * n_own_cells cells of the tables are this rank's cells cell_begin .. cell_begin + n_own_cells - 1; the remaining
   * ncells - n_own_cells cells are GHOSTS -- copies of the other ranks' cells that share a dof with an own cell, their global
   * ids in ghost_cell_ids.  A dof belongs to the rank of its lowest-numbered incident cell; a rank assembles the complete
   * lists of its dofs from its own and ghost cells, so there is NO exchange step (ghost cell matrices are recomputed).
   * Lists / residual entries of dofs the rank does not own are left undefined.  0 / NULL: all cells are own cells. */
  int64_t n_own_cells;
  const int64_t *ghost_cell_ids; /* [ncells - n_own_cells] */
} gm_desc;

/* Timings of the last gm_assemble / gm_assemble_device call, milliseconds, CUDA events on the
 * handle's stream. */
typedef struct gm_times {
  double h2d_ms;      /* iterate down */
  double kernel_ms;   /* all numeric kernels (+ interface exchange) */
  double d2h_ms;      /* nzval / residual up */
  double exchange_ms; /* NCCL interface-row reduction inside kernel_ms */
  int64_t launches;   /* kernels launched by the call */
  int64_t h2d_bytes, d2h_bytes;
} gm_times;

int gm_version(void);
const char *gm_last_error(void);

/* Replaces the construction of the assembler inside FEOperator(res,jac,X,Y) (Solver.jl:157):
 * copies tables to the device, builds constant tables. */
int gm_create(const gm_desc *desc, gm_handle **out);
void gm_destroy(gm_handle *h);

/* Physical parameters captured by the closures at Solver.jl:111-152
 * (Pr, Ra, n, reg=1e-3, g=(0,1), jac_scaling). */
int gm_set_params(gm_handle *h, double Pr, double Ra, double n, double reg, double gx, double gy,
                  double jac_scaling);

/* Replaces allocate_jacobian(op, uh) (the symbolic pass of SparseMatrixAssembler): builds the
 * sparsity pattern and the cell->nnz index map on the device.  *nnz receives the number of
 * stored entries (of the rows/columns owned by this rank). */
int gm_pattern(gm_handle *h, int64_t *nnz);

/* Copies the pattern to the host in Gridap's integer type.  ptr has n_owned+1 entries, idx has
 * nnz entries; both are written with `index_base` (1 for Julia).  For GM_LAYOUT_CSC these are
 * (colptr,rowval) of A, for GM_LAYOUT_CSR (rowptr,colind). */
int gm_get_pattern(gm_handle *h, int64_t *ptr, int64_t *idx, int32_t index_base);

/* Verification hook (bench.py --verify, tests at sizes where the whole matrix cannot be compared): copies
 * the lists of `nrows` selected major dofs (0-based global ids) to the host.  row_ptr[nrows+1] receives
 * offsets into idx_out / val_out (capacity `cap` entries; GM_ERR_ARG if too small), idx_out the minor
 * ids (0-based), val_out the entries of the DEVICE value array d_nzval (NULL: skipped), b_out[nrows] the
 * entries of the DEVICE residual d_resid (NULL: skipped).  Same information as A[:, j] / b[j] in Julia. */
int gm_get_rows(gm_handle *h, int64_t nrows, const int64_t *rows, int64_t *row_ptr, int64_t *idx_out,
                double *val_out, double *b_out, int64_t cap, const double *d_nzval, const double *d_resid);

/* Replaces jacobian!(A,op,uh) / residual!(b,op,uh) / residual_and_jacobian!(b,A,op,uh).
 * x: n_free_total doubles (host).  nzval: nnz doubles or NULL.  resid: n doubles or NULL.
 * Blocks until the outputs are in host memory. */
int gm_assemble(gm_handle *h, const double *x, double *nzval, double *resid, uint32_t flags);

/* Same with device-resident buffers (no host copies): used when the caller keeps the matrix on
 * the GPU, and by bench.py for the HBM-resident figure.  Asynchronous on the handle's stream
 * unless `sync` != 0. */
int gm_assemble_device(gm_handle *h, const double *d_x, double *d_nzval, double *d_resid,
                       uint32_t flags, int32_t sync);

/* Optional: page-lock caller buffers that are reused across Newton iterations (A.nzval, b, x)
 * so the copies in gm_assemble run at full PCIe speed. */
int gm_host_register(void *ptr, int64_t bytes);
int gm_host_unregister(void *ptr);

/* Run the numeric kernels on a caller-provided CUDA stream (cudaStream_t as void*). */
int gm_set_stream(gm_handle *h, void *stream);

int gm_timing(gm_handle *h, gm_times *t);

/* Introspection used by tests/bench: number of owned rows, which numeric path is active
 * (gm_path), number of colours of the scatter path, bytes of device memory held. */
int gm_info(gm_handle *h, int64_t *n_owned, int32_t *active_path, int32_t *ncolors,
            int64_t *device_bytes);

/* Multi-GPU (one process per GPU).  gm_nccl_unique_id writes 128 bytes on rank 0; the caller
 * broadcasts them (torch.distributed / MPI / a file) and every rank calls gm_comm_init.  The
 * interface-row sum of gm_assemble then runs over NCCL. */
int gm_nccl_unique_id(void *id128);
int gm_comm_init(gm_handle *h, const void *id128);

/* Interface buffers of a row block, for callers that move them themselves (tests emulate two ranks
 * on one GPU this way; GM_NO_EXCHANGE).  side 0 = lower interface (sent to rank-1), 1 = upper
 * (received from rank+1).  Buffers are device pointers of gm_iface_count doubles. */
int gm_iface_count(gm_handle *h, int32_t side, int64_t *ndoubles);
int gm_iface_pack(gm_handle *h, const double *d_nzval, const double *d_resid, double *d_buf, uint32_t flags);
int gm_iface_unpack_add(gm_handle *h, double *d_nzval, double *d_resid, const double *d_buf, uint32_t flags);

/*
 * Post-processing assemblies that follow the Newton solve (/root/reference/src/Solver.jl:197-292): the stream function and
 * heat function problems (AffineFEOperator on a P1 space, Solver.jl:200-210 and 254-273) and the entropy fields
 * (interpolate + integrals, Solver.jl:281-292).  `which`: 0 = stream function (Measure(Ω,2)), 1 = heat function (Measure(Ω,4)).
 * Tables are inputs like in gm_desc: in production Gridap supplies them.
 */
typedef struct gm_post_desc {
  int32_t struct_size;
  int32_t nq2, nq4;                     /* points of Measure(Ω,2) and Measure(Ω,4) */
  const int32_t *cell_dofs_psi;         /* [ncells*nv] signed 1-based dof ids of the stream-function space (Dirichlet on V_diri_tags) */
  int64_t n_free_psi, n_diri_psi;
  const double *diri_psi;               /* [n_diri_psi] (0.0 in the reference) */
  const int32_t *cell_dofs_pi;          /* same for the heat-function space (Dirichlet on T_natural_tags, or none) */
  int64_t n_free_pi, n_diri_pi;
  const double *diri_pi;
  /* per rule: weights [nq], vertex (= geometry) basis [nq*nv] and reference gradients [nq*nv*2], P2 basis [nq*nn] and gradients [nq*nn*2] */
  const double *w2, *g2, *dg2, *phi2, *dphi2;
  const double *w4, *g4, *dg4, *phi4, *dphi4;
  const double *dphi_nodes;             /* [nn*nn*2] P2 reference gradients at the nn P2 nodes (entropy interpolation) */
  const double *dg_nodes;               /* [nn*nv*2] geometry gradients at the nodes */
  const int32_t *cell_nodes;            /* [ncells*nn] 0-based node ids of the P2 H1 space the entropies are interpolated into */
  int64_t n_nodes;
} gm_post_desc;

int gm_post_create(gm_handle *h, const gm_post_desc *desc);
/* sizes of problem `which`; pattern of its CSC matrix (colptr n+1, rowval nnz) */
int gm_post_sizes(gm_handle *h, int32_t which, int64_t *n, int64_t *nnz);
int gm_post_get_pattern(gm_handle *h, int32_t which, int64_t *colptr, int64_t *rowval, int32_t index_base);
/* assembles matrix values (nnz) and right-hand side (n) of problem `which` at the solution x (host arrays) */
int gm_post_assemble(gm_handle *h, int32_t which, const double *x, double *nzval, double *b);
/* nodal entropy fields sth, sfl [n_nodes] and totals = {Sth_tot, Sfl_tot, Be_avg} (Solver.jl:281-292) */
int gm_post_entropy(gm_handle *h, const double *x, double *sth, double *sfl, double *totals);

#ifdef __cplusplus
}
#endif
#endif /* GEMOTION_B200_H */
