# SetIntersectionProjectionB200.jl — ccall binding of libsipb200.so for the reference's PARSDMM entry point.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image nor on the GPU box.  The file
# mirrors, call for call, the tested Python/ctypes driver (setintersectionprojection.jl_b200/solver.py) and is
# the stub a maintainer of slimgroup/SetIntersectionProjection.jl would drop next to src/PARSDMM.jl
# (see INTEGRATION.md).  Struct layouts must match include/sipb200.h field for field.
module SIPB200
const lib = joinpath(@__DIR__, "libsipb200.so")

struct SparseOp           # sipb_sparse: custom_TD_OP[1] as CSR (of A) and CSC (colptr/rowval/nzval .- 1), host arrays
    rows::Int64; cols::Int64; nnz::Int64
    rowptr::Ptr{Int64}; colidx::Ptr{Int32}; val::Ptr{Cvoid}
    colptr::Ptr{Int64}; rowidx::Ptr{Int32}; valt::Ptr{Cvoid}
end

struct SetDesc            # sipb_set_desc
    set_kind::Int32; op_kind::Int32; block_mode::Int32; ncvx::Int32
    min::Float64; max::Float64; k::Int64
    min_vec::Ptr{Cvoid}; max_vec::Ptr{Cvoid}
    fiber_axis::Int32; reserved::Int32; td_n::NTuple{3,Int64}     # fiber modes: axis and set_Prop.TD_n[i]
    sparse::Ptr{SparseOp}                                         # custom_TD_OP (op_kind 6), else C_NULL
end

struct Options            # sipb_options
    maxit::Int32; rho_update_frequency::Int32; adjust_rho::Int32; adjust_gamma::Int32
    adjust_feasibility_rho::Int32; zero_ini_guess::Int32; n_rho_ini::Int32; profile_kernels::Int32
    evol_rel_tol::Float64; feas_tol::Float64; obj_tol::Float64; gamma_ini::Float64
    rho_ini::Ptr{Float64}
    fixed_iterations::Int32; return_ly::Int32; resident_io::Int32; warm_resident::Int32
end

mutable struct Log        # sipb_log (arrays are caller-allocated, maxit rows, row-major)
    iters::Int32; feas_rows::Int32; stopped_feasible::Int32; p::Int32; pp::Int32
    set_feasibility::Ptr{Float64}; r_dual::Ptr{Float64}; r_pri::Ptr{Float64}
    r_dual_total::Ptr{Float64}; r_pri_total::Ptr{Float64}; obj::Ptr{Float64}; evol_x::Ptr{Float64}
    rho::Ptr{Float64}; gamma::Ptr{Float64}; cg_it::Ptr{Int32}; cg_relres::Ptr{Float64}
    phase_seconds::NTuple{7,Float64}; solve_seconds::Float64; device_seconds::Float64
    kernel_launches::NTuple{24,Int64}; kernel_ms::NTuple{24,Float64}; kernel_bytes::NTuple{24,Float64}
    total_launches::Int64; h2d_bytes::Int64; d2h_bytes::Int64
end

check(rc) = rc == 0 || error(unsafe_string(ccall((:sipb_last_error, lib), Cstring, ())))

struct DeviceProjector{TF}        # still callable on a CPU vector: P(v) -> sipb_project
    set_kind::Int32; min::Union{TF,Vector{TF}}; max::Union{TF,Vector{TF}}; k::Int64
    fiber_axis::Int32; td_n::NTuple{3,Int64}      # ("fiber","x"|"y"|"z") modes of bounds / cardinality
end

const ctx = Ref{Ptr{Cvoid}}(C_NULL)
function context()
    ctx[] == C_NULL && check(ccall((:sipb_ctx_create, lib), Cint, (Cint, Ptr{Ptr{Cvoid}}), 0, ctx))
    ctx[]
end

op_kind(tag) = Dict("identity"=>0,"D_x"=>1,"D_y"=>2,"D_z"=>3,"TV"=>4,"D2D"=>4,"D3D"=>4,"D_xz"=>5)[tag[2]]

function device_problem(::Type{TF}, AtA, TD_OP, set_Prop, P_sub, comp_grid, options) where TF
    pb = Ref{Ptr{Cvoid}}(C_NULL)
    n = Int64[comp_grid.n...]; h = Float64[TF.(comp_grid.d)...]
    check(ccall((:sipb_problem_create, lib), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Float64}, Cint, Cint, Ptr{Ptr{Cvoid}}),
                context(), TF == Float32 ? 0 : 1, length(n), n, h, options.Minkowski, options.feasibility_only, pb))
    p = length(TD_OP); pp = options.feasibility_only ? p : p - 1
    for i in 1:p
        P = i <= pp ? P_sub[i] : nothing
        d = SetDesc(i <= pp ? P.set_kind : 7, op_kind(set_Prop.tag[i]), 0, set_Prop.ncvx[i],
                    (i <= pp && P.min isa Real) ? P.min : 0.0, (i <= pp && P.max isa Real) ? P.max : 0.0,
                    i <= pp ? P.k : 0,
                    (i <= pp && P.min isa Vector) ? pointer(P.min) : C_NULL,
                    (i <= pp && P.max isa Vector) ? pointer(P.max) : C_NULL,
                    i <= pp ? P.fiber_axis : 0, 0, i <= pp ? P.td_n : (0, 0, 0),
                    C_NULL)    # custom_TD_OP: Ref(SparseOp(...)) built from TD_OP[i] (CSC as stored, CSR = sparse(TD_OP[i]'))
        check(ccall((:sipb_problem_add_set, lib), Cint, (Ptr{Cvoid}, Ref{SetDesc}), pb[], d))
        R = AtA[i]::Matrix{TF}; off = Int64.(set_Prop.AtA_offsets[i])      # exactly what mat2CDS returned
        GC.@preserve R off check(ccall((:sipb_problem_set_ata, lib), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Ptr{Int64}, Cint), pb[], i-1, R, size(R,1), off, length(off)))
    end
    check(ccall((:sipb_problem_finalize, lib), Cint, (Ptr{Cvoid},), pb[]))
    pb[]
end

function PARSDMM(m::Vector{TF}, AtA, TD_OP, set_Prop, P_sub, comp_grid, options,
                 x=zeros(TF,length(m)), l=[], y=[]) where {TF<:Real}
    convert_options!(options, TF)
    options.parallel && error("options.parallel=true is rejected on the device path")
    pb = get!(() -> device_problem(TF, AtA, TD_OP, set_Prop, P_sub, comp_grid, options), PROBLEMS, objectid(AtA))
    p = length(TD_OP); pp = options.feasibility_only ? p : p - 1; maxit = options.maxit
    N = options.Minkowski ? 2length(m) : length(m)
    length(x) == N || (x = [x; zeros(TF, N - length(x))])                       # PARSDMM.jl:85-89
    isempty(l) && (l = [zeros(TF, size(TD_OP[i],1)) for i in 1:p]; y = deepcopy(l))
    rho = Float64.(options.rho_ini)
    o = Options(maxit, options.rho_update_frequency, options.adjust_rho, options.adjust_gamma,
                options.adjust_feasibility_rho, options.zero_ini_guess, length(rho), 0,
                options.evol_rel_tol, options.feas_tol, options.obj_tol, options.gamma_ini, pointer(rho), 0, 1, 0, 0)
    A = Dict(k => zeros(Float64, k == :set_feasibility ? (pp, maxit+2) : (k in (:r_dual,:r_pri,:rho,:gamma) ? (p, maxit) : (maxit,)))
             for k in (:set_feasibility,:r_dual,:r_pri,:r_dual_total,:r_pri_total,:obj,:evol_x,:rho,:gamma,:cg_relres))
    cg_it = zeros(Int32, maxit)                       # (row-major [maxit][p] == column-major (p, maxit))
    lg = Log(0,0,0,0,0, pointer(A[:set_feasibility]), pointer(A[:r_dual]), pointer(A[:r_pri]), pointer(A[:r_dual_total]),
             pointer(A[:r_pri_total]), pointer(A[:obj]), pointer(A[:evol_x]), pointer(A[:rho]), pointer(A[:gamma]),
             pointer(cg_it), pointer(A[:cg_relres]), ntuple(_->0.0,7), 0.0, 0.0, ntuple(_->0,24), ntuple(_->0.0,24), 0,0,0)
    lp = Ptr{Cvoid}[pointer(v) for v in l]; yp = Ptr{Cvoid}[pointer(v) for v in y]
    GC.@preserve m x l y rho A cg_it lp yp check(ccall((:sipb_solve, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ref{Options}, Ref{Log}),
        pb, m, x, lp, yp, o, lg))
    it = max(lg.iters, 1); c = lg.feas_rows
    to = TimerOutput()      # rebuild the seven sections of PARSDMM.jl:40,100,105,113,152,163,229 from lg.phase_seconds
    log_PARSDMM = log_type_PARSDMM(permutedims(A[:set_feasibility][:,1:c]), permutedims(A[:r_dual][:,1:it]),
        permutedims(A[:r_pri][:,1:it]), A[:r_dual_total][1:it], A[:r_pri_total][1:it], A[:obj][1:it], A[:evol_x][1:it],
        permutedims(A[:rho][:,1:it]), permutedims(A[:gamma][:,1:it]), cg_it[1:it], A[:cg_relres][1:it], to)
    return x, log_PARSDMM, l, y
end
end # module
