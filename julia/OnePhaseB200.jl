# OnePhaseB200.jl -- Julia shim that drops libonephase_b200.so in behind OnePhase.jl's own
# plugin points.  `include` it from src/OnePhase.jl after "kkt_system_solver/include.jl"
# (it uses the package's abstract types and helpers), then select it with
#     pars.kkt.kkt_solver_type    = :schur_b200
#     pars.kkt.linear_solver_type = :b200
# or, through JuMP:  set_optimizer_attribute(model, "kkt!kkt_solver_type", :schur_b200)
#                    set_optimizer_attribute(model, "kkt!linear_solver_type", :b200)
#
# UNTESTED IN THIS REPOSITORY'S IMAGE: there is no Julia toolchain here.  The file mirrors,
# call for call, the Python host in onephase.jl_b200/kkt.py, which the GPU parity tests drive.
#
# Reference interfaces replaced (paths under OnePhase.jl/src):
#   linear_solver_B200        <-> linear_solver_JULIA      linear_system_solvers/julia.jl:1-113
#   Schur_B200_KKT_solver     <-> Schur_KKT_solver         kkt_system_solver/schur.jl:3-182
#   ipopt_strategy! (method)  <-> ipopt_strategy!          IPM/delta_strategy.jl:37-114
#   pick_KKT_solver (branch)  <-> pick_KKT_solver          kkt_system_solver/kkt_system_solver.jl:232-287

using SparseArrays, LinearAlgebra, Libdl

const LIBOPB = get(ENV, "ONEPHASE_B200_LIB", "libonephase_b200.so")
const OPB_MODE_CHOLESKY = Cint(0)
const OPB_MODE_LDLT = Cint(1)

mutable struct OpbHandle
    ptr::Ptr{Cvoid}
    function OpbHandle(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:opb_create, LIBOPB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cuint), ref, device, 0)
        h = new(ref[])
        rc == 0 || error("opb_create failed ($rc): " * opb_error(h))
        # finalize! is never called by the IPM (kkt_system_solver.jl:21-25): free HBM from a finalizer
        finalizer(x -> (x.ptr != C_NULL && ccall((:opb_destroy, LIBOPB), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
        return h
    end
end
opb_error(h::OpbHandle) = unsafe_string(ccall((:opb_last_error, LIBOPB), Cstring, (Ptr{Cvoid},), h.ptr))
opb_check(h::OpbHandle, rc) = rc == 0 || error("libonephase_b200 error $rc: " * opb_error(h))
# symbolic analyses (and their device-side maps) are cached per process, up to 8 patterns: a long-running
# session that is done with a family of problems can give the memory back
opb_cache_clear() = ccall((:opb_cache_clear, LIBOPB), Cint, ())

################################################################################
## L1: linear system solver  (abstract_linear_system_solver, linear_system_solvers.jl:11)
################################################################################
mutable struct linear_solver_B200 <: abstract_linear_system_solver
    _h::Union{OpbHandle,Nothing}
    _factor_defined::Bool
    sym::Symbol          # :definite (Cholesky) or :symmetric (LDL' with inertia)
    safe_mode::Bool
    recycle::Bool        # the symbolic analysis is always reused (pattern-hash cache in the library)
    device::Int
    function linear_solver_B200(sym::Symbol, safe_mode::Bool, recycle::Bool)
        (sym == :definite || sym == :symmetric) || error("this.options.sym = $sym not supported")
        return new(nothing, false, sym, safe_mode, recycle, parse(Int, get(ENV, "ONEPHASE_B200_DEVICE", "0")))
    end
end

function initialize!(solver::linear_solver_B200)
    solver._h === nothing && (solver._h = OpbHandle(solver.device))
end
function finalize!(solver::linear_solver_B200)
    solver._h = nothing; solver._factor_defined = false
end

function ls_factor!(solver::linear_solver_B200, SparseMatrix::SparseMatrixCSC{Float64,Int64}, n::Int64, m::Int64, timer::class_advanced_timer)
    start_advanced_timer(timer, "B200/factorize")
    initialize!(solver)
    mode = solver.sym == :definite ? OPB_MODE_CHOLESKY : OPB_MODE_LDLT
    solver.sym == :definite && @assert(m == 0)
    ok = Ref{Cint}(0)
    rc = ccall((:opb_ls_factor_csc, LIBOPB), Cint,
               (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Cint, Int64, Int64, Ref{Cint}),
               solver._h.ptr, size(SparseMatrix, 1), SparseMatrix.colptr, SparseMatrix.rowval, SparseMatrix.nzval,
               1, mode, n, m, ok)
    opb_check(solver._h, rc)
    solver._factor_defined = true
    pause_advanced_timer(timer, "B200/factorize")
    return Int(ok[])        # 1 = inertia (n, m); anything else makes ipopt_strategy! raise delta
end

function ls_solve!(solver::linear_solver_B200, my_rhs::Array{Float64,1}, my_sol::Array{Float64,1}, timer::class_advanced_timer)
    start_advanced_timer(timer, "B200/ls_solve")
    opb_check(solver._h, ccall((:opb_ls_solve, LIBOPB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), solver._h.ptr, my_rhs, my_sol))
    pause_advanced_timer(timer, "B200/ls_solve")
end

function ls_solve(solver::linear_solver_B200, my_rhs::AbstractArray, timer::class_advanced_timer)
    rhs = Vector{Float64}(my_rhs)            # a SparseVector is densified like julia.jl:107-110
    sol = similar(rhs)
    ls_solve!(solver, rhs, sol, timer)
    return sol
end

################################################################################
## L2: KKT system solver  (abstract_schur_solver, kkt_system_solver.jl:13-19)
################################################################################
mutable struct Schur_B200_KKT_solver <: abstract_schur_solver
    ls_solver::abstract_linear_system_solver
    factor_it::Class_iterate
    delta_x_vec::Array{Float64,1}
    delta_s_vec::Array{Float64,1}
    rhs::System_rhs
    dir::Class_point
    kkt_err_norm::Class_kkt_error
    rhs_norm::Float64
    pars::Class_parameters
    schur_diag::Array{Float64,1}
    ready::Symbol
    Q::SparseMatrixCSC{Float64,Int64}      # host copy of tril(Q), refreshed lazily for is_diag_dom
    current_it::Class_iterate
    reduct_factors::Class_reduction_factors
    # B200 state
    _h::Union{OpbHandle,Nothing}
    _pattern::UInt64
    _delta::Float64
    _diag_min::Float64
    _shard_rank::Int                       # one instance over several GPUs (shard_init!); world 1 = off
    _shard_world::Int
    _allgather::Union{Function,Nothing}    # transport of the 384-byte descriptors (shard_init!)
    function Schur_B200_KKT_solver()
        this = new()
        this.ready = :not_ready
        this._h = nothing; this._pattern = 0; this._delta = 0.0; this._diag_min = NaN
        this._shard_rank = 0; this._shard_world = 1; this._allgather = nothing
        return this
    end
end

function initialize!(kkt_solver::Schur_B200_KKT_solver, intial_it::Class_iterate)
    kkt_solver._h === nothing && (kkt_solver._h = OpbHandle(parse(Int, get(ENV, "ONEPHASE_B200_DEVICE", "0"))))
    kkt_solver.dir = zero_point(dim(intial_it), ncon(intial_it))
end

function form_system!(kkt_solver::Schur_B200_KKT_solver, iter::Class_iterate, timer::class_advanced_timer)
    start_advanced_timer(timer, "SCHUR"); start_advanced_timer(timer, "SCHUR/form_system")
    J = get_jac(iter); H = get_lag_hess(iter)      # H lower triangular (eval.jl:132-134)
    h = kkt_solver._h
    pat = hash(J.colptr, hash(J.rowval, hash(H.colptr, hash(H.rowval, UInt64(size(J, 1))))))
    if pat != kkt_solver._pattern
        # first call, or the MOI path dropped numerical zeros (Class_cutest.jl:490-502): new symbolic analysis
        opb_check(h, ccall((:opb_set_structure, LIBOPB), Cint,
                           (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Cint),
                           h.ptr, size(J, 2), size(J, 1), J.colptr, J.rowval, H.colptr, H.rowval, 1))
        kkt_solver._pattern = pat
        # sharded instance: the peer buffers belong to the structure, so the ranks exchange their
        # descriptors again (every rank takes this branch in the same call: same data, same pattern)
        kkt_solver._shard_world > 1 && shard_attach!(kkt_solver, kkt_solver._allgather)
    end
    n = size(J, 2)
    kkt_solver.schur_diag = Vector{Float64}(undef, n)
    dmin = Ref{Float64}(NaN)
    opb_check(h, ccall((:opb_form, LIBOPB), Cint,
                       (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                       h.ptr, J.nzval, H.nzval, get_y(iter), get_s(iter), kkt_solver.schur_diag, dmin))
    kkt_solver._diag_min = dmin[]
    kkt_solver.factor_it = iter
    kkt_solver.ready = :system_formed
    pause_advanced_timer(timer, "SCHUR/form_system"); pause_advanced_timer(timer, "SCHUR")
end

diag_min(kkt_solver::Schur_B200_KKT_solver) = kkt_solver._diag_min

function update_delta_vecs!(kkt_solver::Schur_B200_KKT_solver, delta_x_vec::Array{Float64,1}, delta_s_vec::Array{Float64,1}, timer::class_advanced_timer)
    kkt_solver.delta_x_vec = delta_x_vec
    kkt_solver.delta_s_vec = delta_s_vec
    sum(abs.(delta_s_vec)) > 0.0 && error("Not implemented")       # schur.jl:71
    all(delta_x_vec .== delta_x_vec[1]) || error("Not implemented: non-uniform delta_x_vec")
    kkt_solver._delta = delta_x_vec[1]       # the shift is applied on the device when the fronts are filled
    kkt_solver.ready = :delta_updated
end

function factor_implementation!(kkt_solver::Schur_B200_KKT_solver, timer::class_advanced_timer)
    ok = Ref{Cint}(0)
    opb_check(kkt_solver._h, ccall((:opb_factor, LIBOPB), Cint, (Ptr{Cvoid}, Float64, Ref{Cint}), kkt_solver._h.ptr, kkt_solver._delta, ok))
    return Int(ok[])
end

function compute_direction_implementation!(kkt_solver::Schur_B200_KKT_solver, timer::class_advanced_timer)
    start_advanced_timer(timer, "SCHUR")
    rhs = kkt_solver.rhs; dir = kkt_solver.dir
    n = length(rhs.dual_r); m = length(rhs.primal_r)
    dir.x = Vector{Float64}(undef, n); dir.y = Vector{Float64}(undef, m); dir.s = Vector{Float64}(undef, m)
    err = Vector{Float64}(undef, 6)
    opb_check(kkt_solver._h, ccall((:opb_direction, LIBOPB), Cint,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              kkt_solver._h.ptr, rhs.dual_r, rhs.primal_r, rhs.comp_r, kkt_solver.pars.kkt.ItRefine_Num, dir.x, dir.y, dir.s, err))
    check_for_nan(dir)
    kkt_solver.kkt_err_norm = Class_kkt_error(err[1], err[2], err[3], err[4], err[5], err[6])   # "N err" = ratio
    kkt_solver.rhs_norm = err[5]
    pause_advanced_timer(timer, "SCHUR")
end

# delta_strategy.jl:37-114 specialised on the B200 solver: the probe at delta.zero, the first
# shift and the x8 retries all run on the device; one ccall returns (status, #fac, delta).
function ipopt_strategy!(iter::Class_iterate, kkt_solver::Schur_B200_KKT_solver, pars::Class_parameters, timer::class_advanced_timer)
    d = pars.delta
    delta = Ref{Float64}(0.0); num_fac = Ref{Cint}(0); status = Ref{Cint}(0)
    opb_check(kkt_solver._h, ccall((:opb_factor_delta_loop, LIBOPB), Cint,
              (Ptr{Cvoid}, Float64, Float64, Float64, Float64, Float64, Float64, Float64, Cint, Ref{Float64}, Ref{Cint}, Ref{Cint}),
              kkt_solver._h.ptr, get_delta(iter), d.zero, d.min, d.max, d.start, d.inc, d.dec, 500, delta, num_fac, status))
    n = length(iter.point.x)
    kkt_solver.delta_x_vec = delta[] * ones(n)
    kkt_solver.delta_s_vec = zeros(ncon(iter))
    kkt_solver._delta = delta[]
    kkt_solver.ready = :factored
    status[] == 1 && return :success, Int(num_fac[]), delta[]
    status[] == 0 && return :failure, iter, delta[]
    error("max it")
end

# pick_KKT_solver (kkt_system_solver.jl:232-287): add this branch before the final `else`
#   elseif kkt_solver_type == :schur_b200
#     my_kkt_solver = Schur_B200_KKT_solver()
#     linear_solver_type == :b200 || error("pick a valid solver!")
#     my_kkt_solver.ls_solver = linear_solver_B200(:definite, safe, recycle)
# and, for the symmetric formulation (the reference's generic Symmetric_KKT_solver, symmetric.jl:35-102,
# runs unchanged on top of the L1 plugin: LDL' on the tensor path, inertia (n, m)), one more case in the
# existing `:symmetric` branch (kkt_system_solver.jl:240-249):
#     elseif linear_solver_type == :b200
#       my_kkt_solver.ls_solver = linear_solver_B200(:symmetric, safe, recycle)

# ---------------------------------------------------------------------------------------------
# One instance over several GPUs (include/onephase_b200.h, opb_shard_*): one Julia process per GPU,
# every process making the same calls with the same data.  `allgather(blob)::Vector{Vector{UInt8}}`
# is any transport the host program has (MPI.Allgather, Distributed.jl, a file): only these 384
# bytes per rank travel through it, the numeric data moves between the GPUs inside the kernels.
#   shard_init!(kkt_solver, rank, world, allgather)   after initialize!, before the first form_system!;
#                                                     form_system! then calls shard_attach! after every
#                                                     opb_set_structure, like kkt.py's _attach_peers
# EXPERIMENTAL: like the rest of this file it has never run (no Julia toolchain where it was written); the
# Python host (onephase.jl_b200/kkt.py, DistShard) is the exercised twin of this code path.
function shard_init!(kkt_solver::Schur_B200_KKT_solver, rank::Integer, world::Integer, allgather::Function)
    opb_check(kkt_solver._h, ccall((:opb_shard_init, LIBOPB), Cint, (Ptr{Cvoid}, Cint, Cint),
                                   kkt_solver._h.ptr, rank, world))
    kkt_solver._shard_rank = rank; kkt_solver._shard_world = world; kkt_solver._allgather = allgather
end

function shard_attach!(kkt_solver::Schur_B200_KKT_solver, allgather::Function)
    blob = Vector{UInt8}(undef, 384)
    opb_check(kkt_solver._h, ccall((:opb_shard_export, LIBOPB), Cint, (Ptr{Cvoid}, Ptr{UInt8}),
                                   kkt_solver._h.ptr, blob))
    blobs = allgather(blob)
    for p in 0:(kkt_solver._shard_world - 1)
        p == kkt_solver._shard_rank && continue
        opb_check(kkt_solver._h, ccall((:opb_shard_attach, LIBOPB), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}),
                                       kkt_solver._h.ptr, p, blobs[p + 1]))
    end
    allgather(UInt8[1])          # nobody launches before every rank has attached
end
