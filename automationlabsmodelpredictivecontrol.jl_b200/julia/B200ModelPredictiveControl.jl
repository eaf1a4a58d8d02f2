# B200ModelPredictiveControl.jl -- the reference-side binding of libmpcb200 (include/mpcb200.h).
#
# UNTESTED HERE: neither the build container nor the GPU box has a Julia toolchain (probed: `julia` missing), so this
# file has not been executed.  It is the thin `ccall` shim a maintainer of AutomationLabsModelPredictiveControl.jl
# would add (INTEGRATION.md walks through it); the identical C ABI is exercised from Python ctypes by tests/.
#
# How it plugs into the reference (paths relative to the reference repo):
#   src/types/types.jl:168-192         add    struct b200_solver_def <: AbstractSolvers end
#   src/sub/solver_selection.jl:9-14   add    b200 = b200_solver_def()  to _IMPLEMENTATION_SOLVER_LIST
#   src/sub/design_mpc.jl:84-104       when the solver tag is b200, build a B200Modeler instead of the JuMP model
#   src/main/computation_mpc.jl:17,38  methods of update_initialization! / calculate! for a B200Modeler (below)
module B200ModelPredictiveControl

using LinearAlgebra

const libmpcb200 = get(ENV, "MPCB200_LIB", "libmpcb200.so")

# ---- C structs (field order and types exactly as in include/mpcb200.h) -------------------------------------------
struct MpcbSettings
    eps_abs::Cdouble; eps_rel::Cdouble; eps_prim_inf::Cdouble; rho::Cdouble; rho_eq_scale::Cdouble
    sigma::Cdouble; alpha::Cdouble
    max_iter::Int32; check_every::Int32; device::Int32; kernel::Int32
    ladder_iter::Int32          # rho ladder for state-box rows (0: off), see include/mpcb200.h
    ladder_kappa::Int32
    n_devices::Int32            # 2..8: one handle drives device_ids[1:n_devices] from this process (batch sharded, see include/mpcb200.h)
    device_ids::NTuple{8,Int32}
    cold_init::Int32            # 0 (default): OSQP cold start; 1: cold starts begin at the clipped unconstrained optimum (include/mpcb200.h)
end

struct MpcbLinearDesc
    nx::Int32; nu::Int32; horizon::Int32
    A::Ptr{Cdouble}; B::Ptr{Cdouble}; Q::Ptr{Cdouble}; R::Ptr{Cdouble}; S::Ptr{Cdouble}; P::Ptr{Cdouble}
    umin::Ptr{Cdouble}; umax::Ptr{Cdouble}; xmin::Ptr{Cdouble}; xmax::Ptr{Cdouble}
    state_constraint::Int32; terminal_mode::Int32
end

struct MpcbInfo
    nx::Int32; nu::Int32; horizon::Int32; nz::Int32; mg::Int32; nt::Int32; nt_pad::Int32; kernel::Int32
    device::Int32; sm_count::Int32
    rho::Cdouble; lambda_min::Cdouble; lambda_max::Cdouble
end

struct MpcbBatchIO
    batch::Int64
    x0::Ptr{Cdouble}; xref::Ptr{Cdouble}; uref::Ptr{Cdouble}
    xref_broadcast::Int32; uref_broadcast::Int32
    warm_u::Ptr{Cdouble}; warm_y::Ptr{Cdouble}
    u::Ptr{Cdouble}; e_u::Ptr{Cdouble}; x::Ptr{Cdouble}; e_x::Ptr{Cdouble}; u0::Ptr{Cdouble}
    status::Ptr{Int32}; iters::Ptr{Int32}
    prim_res::Ptr{Cdouble}; dual_res::Ptr{Cdouble}; objective::Ptr{Cdouble}; y::Ptr{Cdouble}
    inner_iters::Ptr{Int32}
end

# nonlinear path (include/mpcb200.h, section "Nonlinear path")
struct MpcbNnDesc
    arch::Int32; activation::Int32; nx::Int32; nu::Int32; n_neurons::Int32; n_hidden::Int32
    W_in::Ptr{Cdouble}; W_hidden::Ptr{Cdouble}; b_hidden::Ptr{Cdouble}; W_out::Ptr{Cdouble}
end

struct MpcbNmpcDesc
    nn::Ptr{MpcbNnDesc}; horizon::Int32
    Q::Ptr{Cdouble}; R::Ptr{Cdouble}; S::Ptr{Cdouble}; P::Ptr{Cdouble}
    umin::Ptr{Cdouble}; umax::Ptr{Cdouble}; xref::Ptr{Cdouble}; uref::Ptr{Cdouble}
    terminal_mode::Int32; state_constraint::Int32
    xmin::Ptr{Cdouble}; xmax::Ptr{Cdouble}
end

struct MpcbNmpcSettings
    qp::MpcbSettings
    sqp_tol::Cdouble; ls_armijo::Cdouble; ls_noise::Cdouble
    sqp_max_iter::Int32; ls_max_halvings::Int32
end

last_error() = unsafe_string(ccall((:mpcb_last_error, libmpcb200), Cstring, ()))
check(rc, what) = rc == 0 || error("$what failed (rc=$rc): $(last_error())")   # no CPU fallback: errors surface here

function default_settings(; kw...)
    s = Ref{MpcbSettings}()
    ccall((:mpcb_default_settings, libmpcb200), Cvoid, (Ref{MpcbSettings},), s)
    d = Dict{Symbol,Any}(n => getfield(s[], n) for n in fieldnames(MpcbSettings))
    for (k, v) in kw
        if k == :devices          # devices = [0, 1, ...]: one handle, several GPUs (kw `mpc_b200_devices`)
            ids = collect(Int32, v)
            1 <= length(ids) <= 8 || error("devices: 1..8 device ordinals")
            d[:n_devices] = Int32(length(ids)); d[:device] = ids[1]
            d[:device_ids] = ntuple(i -> i <= length(ids) ? ids[i] : Int32(0), 8)
        else
            d[k] = v
        end
    end
    return MpcbSettings((convert(fieldtype(MpcbSettings, n), d[n]) for n in fieldnames(MpcbSettings))...)
end

"""
    tune_rho(desc::MpcbLinearDesc, settings::MpcbSettings, X0, xref, uref; candidates = 7, factor = 2.0) -> rho

Step size for this controller chosen on a SAMPLE of the workload (`mpcb_tune_rho`): the batch-wide counterpart of OSQP's per-problem adaptive
rho (the KKT factor is cached once for the whole batch).  Pass the result as `default_settings(rho = ...)`.
"""
function tune_rho(desc::MpcbLinearDesc, settings::MpcbSettings, X0::Matrix{Float64}, xref::Matrix{Float64}, uref::Matrix{Float64}; candidates::Integer = 7, factor::Float64 = 2.0)
    best = Ref{Cdouble}(0.0)
    GC.@preserve X0 xref uref begin
        io = MpcbBatchIO(size(X0, 2), pointer(X0), pointer(xref), pointer(uref), size(xref, 2) == 1 ? 1 : 0, size(uref, 2) == 1 ? 1 : 0,
                         C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)
        check(ccall((:mpcb_tune_rho, libmpcb200), Cint, (Ref{MpcbLinearDesc}, Ref{MpcbSettings}, Ref{MpcbBatchIO}, Int32, Cdouble, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                    desc, settings, io, Int32(candidates), factor, best, C_NULL, C_NULL), "mpcb_tune_rho")
    end
    return best[]
end

# ---- what `tuning.modeler` holds for mpc_solver = "b200" (types.jl:115 leaves the field untyped) ------------------
mutable struct B200Modeler
    handle::Ptr{Cvoid}
    info::MpcbInfo
    X0::Matrix{Float64}        # nx x batch, set by update_initialization!
    xref::Matrix{Float64}      # nx x (1 | batch)
    uref::Matrix{Float64}      # nu x (1 | batch)
    # batched results (reference layout per problem, problems along the last dimension)
    u::Array{Float64,3}; e_u::Array{Float64,3}; x::Array{Float64,3}; e_x::Array{Float64,3}
    status::Vector{Int32}; iterations::Vector{Int32}
    prim_res::Vector{Float64}; dual_res::Vector{Float64}; objective::Vector{Float64}
end

"""
    B200Modeler(A, B, Q, R, S, P, umin, umax, xmin, xmax, horizon; state_constraint, terminal, settings)

Replaces `_model_predictive_control_modeler_implementation(::LinearProgramming, system, ...)` (linear.jl:20-103) +
`_create_terminal_ingredient` constraints (design_mpc.jl:330-331) + `_create_quadratic_cost_function`
(design_mpc.jl:405-468): the same data, condensed and factored once, cached on the GPU.
"""
function B200Modeler(A, B, Q, R, S, P, umin, umax, xmin, xmax, horizon::Int;
                     state_constraint::Bool=false, terminal::String="none", settings::MpcbSettings=default_settings())
    terminal in ("none", "equality", "contractive") || error("mpc_solver=\"b200\" supports mpc_terminal_ingredient \"none\", \"equality\" and \"contractive\" only")
    nx, nu = size(B)
    mats = map(M -> Matrix{Float64}(M), (A, B, Q, R, S, P))
    vecs = map(v -> Vector{Float64}(v), (umin, umax, xmin, xmax))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve mats vecs begin
        d = MpcbLinearDesc(nx, nu, horizon, map(pointer, mats)..., map(pointer, vecs)..., state_constraint ? 1 : 0,
                           terminal == "equality" ? 1 : terminal == "contractive" ? 2 : 0)
        check(ccall((:mpcb_create_linear, libmpcb200), Cint, (Ref{MpcbLinearDesc}, Ref{MpcbSettings}, Ref{Ptr{Cvoid}}), d, settings, h),
              "mpcb_create_linear")
    end
    info = Ref{MpcbInfo}()
    check(ccall((:mpcb_get_info, libmpcb200), Cint, (Ptr{Cvoid}, Ref{MpcbInfo}), h[], info), "mpcb_get_info")
    m = B200Modeler(h[], info[], zeros(nx, 0), zeros(nx, 1), zeros(nu, 1), zeros(nu, horizon, 0), zeros(nu, horizon, 0),
                    zeros(nx, horizon + 1, 0), zeros(nx, horizon + 1, 0), Int32[], Int32[], Float64[], Float64[], Float64[])
    finalizer(m -> ccall((:mpcb_destroy, libmpcb200), Cvoid, (Ptr{Cvoid},), m.handle), m)
    return m
end

# ---- the hot path: methods the reference's two compute functions gain (computation_mpc.jl:17-55) ------------------
"update_initialization!(C, x0::Vector): single problem, reference semantics (JuMP.fix of x[:,1] becomes the x0 operand)."
function update_initialization!(m::B200Modeler, initialization::AbstractVector, references)
    m.X0 = reshape(Vector{Float64}(initialization), :, 1)
    m.xref = reshape(Vector{Float64}(references.x[:, 1]), :, 1)
    m.uref = reshape(Vector{Float64}(references.u[:, 1]), :, 1)
end

"Batched overload: X0 is nx x batch; xref / uref are nx x 1 (broadcast) or nx x batch."
function update_initialization!(m::B200Modeler, X0::AbstractMatrix; xref::AbstractMatrix, uref::AbstractMatrix)
    m.X0 = Matrix{Float64}(X0); m.xref = Matrix{Float64}(xref); m.uref = Matrix{Float64}(uref)
end

"calculate!(C): one batched solve; fills the batched result arrays (and, in the reference wrapper, C.computation_results)."
function calculate!(m::B200Modeler)
    nx, nu, H = Int(m.info.nx), Int(m.info.nu), Int(m.info.horizon)
    batch = size(m.X0, 2)
    batch > 0 || error("calculate!: call update_initialization! first")
    m.u = Array{Float64}(undef, nu, H, batch); m.e_u = similar(m.u)
    m.x = Array{Float64}(undef, nx, H + 1, batch); m.e_x = similar(m.x)
    m.status = Vector{Int32}(undef, batch); m.iterations = Vector{Int32}(undef, batch)
    m.prim_res = Vector{Float64}(undef, batch); m.dual_res = similar(m.prim_res); m.objective = similar(m.prim_res)
    GC.@preserve m begin
        io = MpcbBatchIO(batch, pointer(m.X0), pointer(m.xref), pointer(m.uref), size(m.xref, 2) == 1 ? 1 : 0,
                         size(m.uref, 2) == 1 ? 1 : 0, C_NULL, C_NULL, pointer(m.u), pointer(m.e_u), pointer(m.x), pointer(m.e_x),
                         C_NULL, pointer(m.status), pointer(m.iterations), pointer(m.prim_res), pointer(m.dual_res),
                         pointer(m.objective), C_NULL, C_NULL)
        check(ccall((:mpcb_solve_linear_batch, libmpcb200), Cint, (Ptr{Cvoid}, Ref{MpcbBatchIO}), m.handle, io), "mpcb_solve_linear_batch")
    end
    return m
end

# ---- nonlinear path: Flux chain -> mpcb_nn_desc, NL modeler + Ipopt -> mpcb_create_nmpc / mpcb_solve_nmpc_batch -----
const _ACTIVATION_IDS = Dict("relu" => 0, "tanh" => 1, "sigmoid" => 2, "σ" => 2, "swish" => 3, "identity" => 4)

"""
    nn_arrays(params, arch, activation)

`params = collect(Flux.params(system.f))` parsed exactly as the reference's NL modelers do (fnn.jl:88-107): params[1] = W_in,
then (W_j, b_j) pairs, params[end] = W_out.  Returns Float64 copies (Flux stores Float32) and the descriptor fields.
"""
function nn_arrays(params::Vector, arch::Symbol, activation::String)
    W_in = Matrix{Float64}(params[1]); W_out = Matrix{Float64}(params[end])
    n_hidden = (length(params) - 2) ÷ 2
    W_h = n_hidden == 0 ? zeros(1) : vcat((vec(Matrix{Float64}(params[i])) for i in 2:2:length(params)-1)...)
    b_h = n_hidden == 0 ? zeros(1) : vcat((Vector{Float64}(params[i+1]) for i in 2:2:length(params)-1)...)
    nx = size(W_out, 1); nu = size(W_in, 2) - nx
    # :icnn and :rbf take the :fnn path: the reference's modelers for them are the Fnn ones up to the type tag (icnn.jl, rbf.jl:61-186)
    arch in (:fnn, :icnn, :rbf, :resnet, :polynet, :densenet) || error("mpc_solver=\"b200\": unsupported network family $arch")
    return (W_in, W_h, b_h, W_out), (Int32(arch == :resnet ? 1 : arch == :polynet ? 2 : arch == :densenet ? 3 : 0), Int32(_ACTIVATION_IDS[activation]), Int32(nx), Int32(nu),
                                       Int32(size(W_in, 1)), Int32(n_hidden))
end

mutable struct B200NonlinearModeler
    handle::Ptr{Cvoid}
    nx::Int; nu::Int; horizon::Int
    X0::Matrix{Float64}; xref::Matrix{Float64}; uref::Matrix{Float64}
    u::Array{Float64,3}; e_u::Array{Float64,3}; x::Array{Float64,3}; e_x::Array{Float64,3}
    status::Vector{Int32}; iterations::Vector{Int32}; inner_iterations::Vector{Int32}
    step::Vector{Float64}; objective::Vector{Float64}
end

"Replaces `_model_predictive_control_modeler_implementation(::NonLinearProgramming, ::Fnn | ::ResNet, ...)` (fnn.jl:63-189,
resnet.jl:62-188) + `_JuMP_model_definition(::NonLinearProgramming, ::ipopt_solver_def)`."
function B200NonlinearModeler(params::Vector, arch::Symbol, activation::String, Q, R, S, P, umin, umax, horizon::Int, xref, uref;
                              terminal::String="none", state_constraint::Bool=false, xmin=zeros(0), xmax=zeros(0),
                              settings::Union{Nothing,MpcbNmpcSettings}=nothing)
    terminal in ("none", "equality", "contractive") || error("mpc_solver=\"b200\" supports mpc_terminal_ingredient \"none\", \"equality\" and \"contractive\" only")
    arrs, f = nn_arrays(params, arch, activation)
    mats = map(M -> Matrix{Float64}(M), (Q, R, S, P)); vecs = map(v -> Vector{Float64}(v), (umin, umax, xref, uref))
    xb = (Vector{Float64}(xmin), Vector{Float64}(xmax))
    st = Ref{MpcbNmpcSettings}()
    settings === nothing ? ccall((:mpcb_default_nmpc_settings, libmpcb200), Cvoid, (Ref{MpcbNmpcSettings},), st) : (st[] = settings)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve arrs mats vecs xb begin
        nd = Ref(MpcbNnDesc(f..., map(pointer, arrs)...))
        GC.@preserve nd begin
            d = MpcbNmpcDesc(Base.unsafe_convert(Ptr{MpcbNnDesc}, nd), horizon, map(pointer, mats)..., map(pointer, vecs)...,
                             terminal == "equality" ? 1 : terminal == "contractive" ? 2 : 0, state_constraint ? 1 : 0,
                             state_constraint ? pointer(xb[1]) : C_NULL, state_constraint ? pointer(xb[2]) : C_NULL)
            check(ccall((:mpcb_create_nmpc, libmpcb200), Cint, (Ref{MpcbNmpcDesc}, Ref{MpcbNmpcSettings}, Ref{Ptr{Cvoid}}), d, st, h), "mpcb_create_nmpc")
        end
    end
    nx, nu = Int(f[3]), Int(f[4])
    m = B200NonlinearModeler(h[], nx, nu, horizon, zeros(nx, 0), zeros(nx, 1), zeros(nu, 1), zeros(nu, horizon, 0), zeros(nu, horizon, 0),
                             zeros(nx, horizon + 1, 0), zeros(nx, horizon + 1, 0), Int32[], Int32[], Int32[], Float64[], Float64[])
    finalizer(m -> ccall((:mpcb_destroy_nmpc, libmpcb200), Cvoid, (Ptr{Cvoid},), m.handle), m)
    return m
end

function update_initialization!(m::B200NonlinearModeler, X0::AbstractMatrix; xref::AbstractMatrix, uref::AbstractMatrix)
    m.X0 = Matrix{Float64}(X0); m.xref = Matrix{Float64}(xref); m.uref = Matrix{Float64}(uref)
end

"method = :non_linear: the SQP solve of the NL modeler; method = :linear: the linear method on the black-box model re-designed per
problem at that problem's reference (design_mpc.jl:319-327) -- mpcb_solve_relinearized_batch."
function calculate!(m::B200NonlinearModeler; method::Symbol = :non_linear)
    nx, nu, H = m.nx, m.nu, m.horizon
    batch = size(m.X0, 2)
    batch > 0 || error("calculate!: call update_initialization! first")
    m.u = Array{Float64}(undef, nu, H, batch); m.e_u = similar(m.u); m.x = Array{Float64}(undef, nx, H + 1, batch); m.e_x = similar(m.x)
    m.status = Vector{Int32}(undef, batch); m.iterations = similar(m.status); m.inner_iterations = similar(m.status)
    m.step = Vector{Float64}(undef, batch); m.objective = similar(m.step)
    GC.@preserve m begin
        io = MpcbBatchIO(batch, pointer(m.X0), pointer(m.xref), pointer(m.uref), size(m.xref, 2) == 1 ? 1 : 0, size(m.uref, 2) == 1 ? 1 : 0,
                         C_NULL, C_NULL, pointer(m.u), pointer(m.e_u), pointer(m.x), pointer(m.e_x), C_NULL, pointer(m.status),
                         pointer(m.iterations), pointer(m.step), C_NULL, pointer(m.objective), C_NULL, pointer(m.inner_iterations))
        if method == :linear
            check(ccall((:mpcb_solve_relinearized_batch, libmpcb200), Cint, (Ptr{Cvoid}, Ref{MpcbBatchIO}), m.handle, io), "mpcb_solve_relinearized_batch")
        else
            check(ccall((:mpcb_solve_nmpc_batch, libmpcb200), Cint, (Ptr{Cvoid}, Ref{MpcbBatchIO}), m.handle, io), "mpcb_solve_nmpc_batch")
        end
    end
    return m
end

"ControlSystems.are(Discrete, A_i, B_i, Q, R) (design_mpc.jl:327) for many systems on the GPU: A nx x nx x n, B nx x nu x n."
function dare_batch(A::Array{Float64,3}, B::Array{Float64,3}, Q::Matrix{Float64}, R::Matrix{Float64}; device::Integer=0)
    nx, nu, n = size(B)
    P = Array{Float64}(undef, nx, nx, n); steps = Vector{Int32}(undef, n)
    check(ccall((:mpcb_dare_batch, libmpcb200), Cint, (Int32, Int64, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}),
                device, n, nx, nu, A, B, Q, R, P, steps), "mpcb_dare_batch")
    return P, steps
end

"AutomationLabsSystems.proceed_system_linearization for a Flux chain (fnn.jl:42, design_mpc.jl:319-323) on the GPU."
function linearize(params::Vector, arch::Symbol, activation::String, x::Vector{Float64}, u::Vector{Float64}; device::Integer=0)
    arrs, f = nn_arrays(params, arch, activation)
    nx, nu = Int(f[3]), Int(f[4])
    A = Matrix{Float64}(undef, nx, nx); B = Matrix{Float64}(undef, nx, nu)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve arrs begin
        nd = MpcbNnDesc(f..., map(pointer, arrs)...)
        check(ccall((:mpcb_create_nn, libmpcb200), Cint, (Ref{MpcbNnDesc}, Int32, Ref{Ptr{Cvoid}}), nd, device, h), "mpcb_create_nn")
    end
    rc = ccall((:mpcb_nn_jacobian_batch, libmpcb200), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
               h[], 1, x, u, C_NULL, A, B)
    ccall((:mpcb_destroy_nn, libmpcb200), Cvoid, (Ptr{Cvoid},), h[])
    check(rc, "mpcb_nn_jacobian_batch")
    return A, B
end

# In AutomationLabsModelPredictiveControl (computation_mpc.jl) the two new methods read:
#
#   function update_initialization!(C::ModelPredictiveControlController, initialization::Vector)
#       C.initialization = initialization
#       if C.tuning.modeler isa B200Modeler
#           return B200ModelPredictiveControl.update_initialization!(C.tuning.modeler, initialization, C.tuning.reference)
#       end
#       ...                                   # existing JuMP.fix loop, unchanged
#   end
#
#   function calculate!(C::ModelPredictiveControlController)
#       if C.tuning.modeler isa B200Modeler
#           m = B200ModelPredictiveControl.calculate!(C.tuning.modeler)
#           C.computation_results.u[:, :]   = m.u[:, :, 1];   C.computation_results.e_u[:, :] = m.e_u[:, :, 1]
#           C.computation_results.x[:, :]   = m.x[:, :, 1];   C.computation_results.e_x[:, :] = m.e_x[:, :, 1]
#           return
#       end
#       ...                                   # existing JuMP.optimize! path, unchanged
#   end

end # module
