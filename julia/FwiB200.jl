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
