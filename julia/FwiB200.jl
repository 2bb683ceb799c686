# FwiB200.jl -- drop-in binding of libfwi_b200.so for FwiFlow.jl (replaces the bodies of `fwi_op` / `fwi_obs_op`,
# src/Core.jl:20-53).  Add `include("FwiB200.jl")` to src/FwiFlow.jl and point LIBFWI at the built library.
#
# UNVERIFIED in the build image of this repository (no Julia toolchain there): the same entry points, argument order
# and layouts are exercised through Python ctypes by tests/ (fwiflow/jl_b200/ops.py is the line-by-line equivalent).
# See INTEGRATION.md for the conventions (row-major (nz, nx) doubles, 0-based Int32 shot ids, MPa, file formats).

const LIBFWI = joinpath(@__DIR__, "../deps/fwi_b200/libfwi_b200.so")

rowmajor(a::AbstractMatrix{Float64}) = permutedims(a)      # (nz,nx) col-major -> [z][x] row-major bytes
fwi_error(rc) = rc == 0 ? nothing :
    error("fwi_b200 ($rc): " * unsafe_string(ccall((:fwi_b200_last_error, LIBFWI), Cstring, ())))

"what the parameter file says: [nz, nx, nSteps, nPoints_pml, nPad, if_win, scratch, 0] (host-only, fwi_b200_para_info)"
function para_info_b200(para::String)
    out = zeros(Int32, 8)
    rc = ccall((:fwi_b200_para_info, LIBFWI), Cint, (Cstring, Ptr{Cint}), para, out)
    fwi_error(rc); out
end

"the C ABI carries no sizes (like the reference's cufd): refuse arrays that do not match the parameter file"
function check_shapes_b200(λ, μ, ρ, stf, shot_ids, para::String)
    info = para_info_b200(para); nz, nx, nsteps = info[1], info[2], info[3]
    for (name, a) in (("lambda", λ), ("mu", μ), ("den", ρ))
        size(a) == (nz, nx) || error("fwi_b200: $name has size $(size(a)); the parameter file says ($nz, $nx)")
    end
    size(stf, 2) == nsteps || error("fwi_b200: stf has $(size(stf, 2)) samples per row; the parameter file says $nsteps")
    (isempty(shot_ids) || minimum(shot_ids) < 0 || maximum(shot_ids) >= size(stf, 1)) &&
        error("fwi_b200: shot ids do not index the $(size(stf, 1)) rows of stf (row = global shot id)")
    nothing
end

"loss = fwi_op_b200(λ, μ, ρ, stf, gpu_id, shot_ids0, para)   (calc_id 0; FwiOp.cpp:47-103)"
function fwi_op_b200(λ, μ, ρ, stf, gpu_id::Integer, shot_ids::Vector{Int32}, para::String)
    check_shapes_b200(λ, μ, ρ, stf, shot_ids, para)
    misfit = Ref{Cdouble}(0.0)
    rc = ccall((:fwi_b200_forward, LIBFWI), Cint,
               (Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cint}, Cstring),
               misfit, rowmajor(λ), rowmajor(μ), rowmajor(ρ), rowmajor(stf), gpu_id, length(shot_ids), shot_ids, para)
    fwi_error(rc); misfit[]
end

"(gλ, gμ, gρ, g_stf) = fwi_op_grad_b200(...)                  (calc_id 1; FwiOp.cpp:130-223)"
function fwi_op_grad_b200(λ, μ, ρ, stf, gpu_id::Integer, shot_ids::Vector{Int32}, para::String)
    check_shapes_b200(λ, μ, ρ, stf, shot_ids, para)
    nz, nx = size(λ); nsteps = size(stf, 2)
    gλ = zeros(nx, nz); gμ = zeros(nx, nz); gρ = zeros(nx, nz)          # row-major (nz,nx) buffers
    gs = zeros(nsteps, length(shot_ids))                                # row-major (group, nSteps)
    rc = ccall((:fwi_b200_backward, LIBFWI), Cint,
               (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                Ptr{Cdouble}, Cint, Cint, Ptr{Cint}, Cstring),
               gλ, gμ, gρ, gs, rowmajor(λ), rowmajor(μ), rowmajor(ρ), rowmajor(stf),
               gpu_id, length(shot_ids), shot_ids, para)
    fwi_error(rc)
    permutedims(gλ), permutedims(gμ), permutedims(gρ), permutedims(gs)
end

"writes <data_dir>/Shot<id>.bin                               (calc_id 2; FwiOp.cpp:260-317)"
function fwi_obs_op_b200(λ, μ, ρ, stf, gpu_id::Integer, shot_ids::Vector{Int32}, para::String)
    check_shapes_b200(λ, μ, ρ, stf, shot_ids, para)
    misfit = Ref{Cdouble}(0.0)
    rc = ccall((:fwi_b200_obscalc, LIBFWI), Cint,
               (Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cint}, Cstring),
               misfit, rowmajor(λ), rowmajor(μ), rowmajor(ρ), rowmajor(stf), gpu_id, length(shot_ids), shot_ids, para)
    fwi_error(rc); misfit[]
end

"(loss, gλ, gμ, gρ, g_stf) on several GPUs of this process"
function fwi_op_and_grad_multi_b200(λ, μ, ρ, stf, gpu_ids::Vector{Int32}, shot_ids::Vector{Int32}, para::String)
    check_shapes_b200(λ, μ, ρ, stf, shot_ids, para)
    nz, nx = size(λ); nsteps = size(stf, 2)
    misfit = Ref{Cdouble}(0.0)
    gλ = zeros(nx, nz); gμ = zeros(nx, nz); gρ = zeros(nx, nz); gs = zeros(nsteps, length(shot_ids))
    rc = ccall((:fwi_b200_gradient_multi, LIBFWI), Cint,
               (Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Cstring),
               misfit, gλ, gμ, gρ, gs, rowmajor(λ), rowmajor(μ), rowmajor(ρ), rowmajor(stf),
               length(gpu_ids), gpu_ids, length(shot_ids), shot_ids, para)
    fwi_error(rc)
    misfit[], permutedims(gλ), permutedims(gμ), permutedims(gρ), permutedims(gs)
end

# ---- the reference's operator surface, unchanged (src/Core.jl:20-53; shot ids stay 0-based here as there) ----------
"misfit = fwi_op(λ, μ, ρ, stf, gpu_id, shot_ids, para_fname)"
fwi_op(λ::Array{Float64}, μ::Array{Float64}, ρ::Array{Float64}, stf::Array{Float64}, gpu_id::Integer,
       shot_ids::Array{<:Integer}, para_fname::String) =
    fwi_op_b200(λ, μ, ρ, stf, gpu_id, convert(Vector{Int32}, vec(shot_ids)), para_fname)

"writes Data/Shot<id>.bin; returns 0.0 like the reference's op"
fwi_obs_op(λ::Array{Float64}, μ::Array{Float64}, ρ::Array{Float64}, stf::Array{Float64}, gpu_id::Integer,
           shot_ids::Array{<:Integer}, para_fname::String) =
    fwi_obs_op_b200(λ, μ, ρ, stf, gpu_id, convert(Vector{Int32}, vec(shot_ids)), para_fname)

"(misfit, gλ, gμ, gρ, g_stf) from ONE forward propagation (fwi_b200_misfit_and_gradient)"
function fwi_op_and_grad_b200(λ, μ, ρ, stf, gpu_id::Integer, shot_ids::Vector{Int32}, para::String)
    check_shapes_b200(λ, μ, ρ, stf, shot_ids, para)
    nz, nx = size(λ); nsteps = size(stf, 2)
    misfit = Ref{Cdouble}(0.0)
    gλ = zeros(nx, nz); gμ = zeros(nx, nz); gρ = zeros(nx, nz); gs = zeros(nsteps, length(shot_ids))
    rc = ccall((:fwi_b200_misfit_and_gradient, LIBFWI), Cint,
               (Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cint}, Cstring),
               misfit, gλ, gμ, gρ, gs, rowmajor(λ), rowmajor(μ), rowmajor(ρ), rowmajor(stf),
               gpu_id, length(shot_ids), shot_ids, para)
    fwi_error(rc)
    misfit[], permutedims(gλ), permutedims(gμ), permutedims(gρ), permutedims(gs)
end

"free the device contexts cached behind the host-buffer entry points"
fwi_release_b200() = ccall((:fwi_b200_release, LIBFWI), Cvoid, ())

# ---- without any transpose: Julia matrices are column-major, which is the order the device keeps (layout = 1) ----------
"(misfit, gλ, gμ, gρ, g_stf) = cufd_colmajor_b200(calc_id, λ, μ, ρ, stf, gpu_id, shot_ids0, para): the (nz, nx) arrays go
down and come back as they are (fwi_b200_cufd_ex, layout = 1); the reference transposes twice per call
(convert_to_tensor + libCUFD.cu:68-78)"
function cufd_colmajor_b200(calc_id::Integer, λ::Matrix{Float64}, μ::Matrix{Float64}, ρ::Matrix{Float64}, stf,
                            gpu_id::Integer, shot_ids::Vector{Int32}, para::String)
    check_shapes_b200(λ, μ, ρ, stf, shot_ids, para)
    nz, nx = size(λ); nsteps = size(stf, 2)
    misfit = Ref{Cdouble}(0.0)
    gλ = zeros(nz, nx); gμ = zeros(nz, nx); gρ = zeros(nz, nx); gs = zeros(nsteps, length(shot_ids))
    rc = ccall((:fwi_b200_cufd_ex, LIBFWI), Cint,
               (Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cint}, Cstring, Cint, Cint),
               misfit, gλ, gμ, gρ, gs, λ, μ, ρ, rowmajor(stf), calc_id, gpu_id, length(shot_ids), shot_ids, para, 1, 1)
    fwi_error(rc)
    misfit[], gλ, gμ, gρ, permutedims(gs)
end

# ---- resident plan: velocities in, velocity gradients out (what compute_loss_with_stf + gradients(loss, cp) do in
# TensorFlow around the op, src/FWI.jl:156-205, src/Utils.jl:221-227, on the device) ------------------------------------
"device-resident evaluation context of one (para file, gpu, shot group): observations, source functions, wavefield and
frame buffers stay on the GPU between L-BFGS evaluations"
mutable struct PlanB200
    h::Ptr{Cvoid}
    nz::Int; nx::Int; nsteps::Int
end

function PlanB200(para::String, gpu_id::Integer, shot_ids::Vector{Int32}, stf::AbstractMatrix{Float64})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    fwi_error(ccall((:fwi_b200_plan_create, LIBFWI), Cint, (Ref{Ptr{Cvoid}}, Cstring, Cint, Cint, Ptr{Cint}, Cint),
                    h, para, gpu_id, length(shot_ids), shot_ids, 0))
    info = para_info_b200(para)
    p = PlanB200(h[], info[1], info[2], info[3])
    finalizer(close_b200, p)
    fwi_error(ccall((:fwi_b200_plan_set_layout, LIBFWI), Cint, (Ptr{Cvoid}, Cint), p.h, 1))   # Julia matrices as they are
    fwi_error(ccall((:fwi_b200_plan_load_obs_files, LIBFWI), Cint, (Ptr{Cvoid},), p.h))        # Data/Shot<id>.bin, once
    size(stf, 2) == p.nsteps || error("fwi_b200: stf has $(size(stf, 2)) samples per row; the parameter file says $(p.nsteps)")
    fwi_error(ccall((:fwi_b200_plan_set_stf, LIBFWI), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p.h, rowmajor(stf)))
    p
end

function close_b200(p::PlanB200)
    p.h == C_NULL || ccall((:fwi_b200_plan_destroy, LIBFWI), Cvoid, (Ptr{Cvoid},), p.h)
    p.h = C_NULL
    nothing
end

"(misfit, g_cp, g_cs, g_ρ) on the padded grid for cp, cs, ρ given unpadded (nz - 2 nPml - nPad, nx - 2 nPml) or padded
(nz, nx); refs = (cp_ref, cs_ref, ρ_ref) of the same size unless is_masked (src/FWI.jl:174-176)"
function misfit_and_gradient_b200(p::PlanB200, cp::Matrix{Float64}, cs::Matrix{Float64}, ρ::Matrix{Float64};
                                  refs = nothing, is_masked::Bool = false)
    padded = size(cp) == (p.nz, p.nx)
    size(cs) == size(cp) == size(ρ) || error("fwi_b200: cp, cs, ρ differ in size")
    is_masked || refs !== nothing || error("fwi_b200: refs = (cp_ref, cs_ref, ρ_ref) is required unless is_masked")
    r = is_masked ? (C_NULL, C_NULL, C_NULL) : map(a -> pointer(a), refs)
    GC.@preserve refs begin
        fwi_error(ccall((:fwi_b200_plan_set_velocities, LIBFWI), Cint,
                        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint),
                        p.h, cp, cs, ρ, r[1], r[2], r[3], is_masked ? 1 : 0, padded ? 1 : 0))
    end
    fwi_error(ccall((:fwi_b200_plan_run, LIBFWI), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint), p.h, 1, C_NULL, 1))
    misfit = Ref{Cdouble}(0.0)
    g_cp = zeros(p.nz, p.nx); g_cs = zeros(p.nz, p.nx); g_ρ = zeros(p.nz, p.nx)
    fwi_error(ccall((:fwi_b200_plan_get_velocity_gradients, LIBFWI), Cint,
                    (Ptr{Cvoid}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), p.h, misfit, g_cp, g_cs, g_ρ))
    misfit[], g_cp, g_cs, g_ρ
end
