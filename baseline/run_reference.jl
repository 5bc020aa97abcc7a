# run_reference.jl -- dump (Q, rhs, delta, inertia, direction) of every ls_factor!/direction of the
# REAL reference (Julia + CHOLMOD) so that tests/ can pin the oracle and the CUDA path on it.
# Cannot run in this image (no Julia, no network); kept so it can be run the day a toolchain exists:
#   julia --project=/path/to/OnePhase.jl baseline/run_reference.jl out_dir
# It wraps linear_solver_JULIA (src/linear_system_solvers/julia.jl:21-113) without changing it.
using OnePhase, SparseArrays, DelimitedFiles, JuMP
const OUT = length(ARGS) > 0 ? ARGS[1] : "reference_dump"
mkpath(OUT)
const COUNTER = Ref(0)
function dump_csc(name, A::SparseMatrixCSC)
    writedlm(joinpath(OUT, name * "_colptr.txt"), A.colptr); writedlm(joinpath(OUT, name * "_rowval.txt"), A.rowval)
    writedlm(joinpath(OUT, name * "_nzval.txt"), A.nzval)
end
# record every factorisation and solve through method wrappers
const _ls_factor = OnePhase.ls_factor!
function OnePhase.ls_factor!(s::OnePhase.linear_solver_JULIA, Q::SparseMatrixCSC{Float64,Int64}, n::Int64, m::Int64, timer)
    COUNTER[] += 1
    dump_csc("fac$(COUNTER[])_Q", Q)
    inertia = invoke(_ls_factor, Tuple{OnePhase.linear_solver_JULIA,SparseMatrixCSC{Float64,Int64},Int64,Int64,Any}, s, Q, n, m, timer)
    writedlm(joinpath(OUT, "fac$(COUNTER[])_inertia.txt"), [inertia])
    return inertia
end
model = Model(OnePhase.OnePhaseSolver)
@variable(model, x, start = -3)
@objective(model, Min, x)
@NLconstraint(model, x^2 >= 1.0)
@NLconstraint(model, x >= -1.0)
optimize!(model)            # README.md:34-45
