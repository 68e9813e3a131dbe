# TensorKrylovB200.jl -- the reference-side binding of libtensorkrylov_b200.so.
#
# Drop-in for `tensorkrylov!` (TensorKrylov.jl src/tensor_krylov_method.jl:36-125): same signature, same
# ConvergenceData conventions, same three exits.  Load it after `using TensorKrylov`; it adds `tensorkrylov_b200!`
# and `solve_tensorized_system_b200`, and `TensorKrylovB200.install!()` re-points the package's own
# `tensorkrylov!` (hence `solve_tensorized_system` and every driver) at the library for Float64 data.
#
# This file cannot be executed in the build image (no Julia there); every symbol it binds is exercised through
# ctypes by the test-suite (tests/test_host_cpu.py::test_cabi_exports_every_declared_symbol and the -m gpu tests).
module TensorKrylovB200

using TensorKrylov
using TensorKrylov: KronMat, KronProd, ConvergenceData, KruskalTensor, TensorDecomposition,
                    TensorLanczos, TensorLanczosReorth, TensorArnoldi, SymInstance, NonSymInstance,
                    MatrixGallery, LaplaceDense, Laplace, ConvDiff, EigValMat, RandSPD,
                    SpectralData, ApproximationData, update_data!, dimensions
using SparseArrays, LinearAlgebra

const libtk = get(ENV, "TENSORKRYLOV_B200_LIB", "libtensorkrylov_b200.so")

const TK_FLAG_REFERENCE_H1 = Cint(1)

instance_code(::Type{SymInstance})    = Cint(0)
instance_code(::Type{NonSymInstance}) = Cint(1)
class_code(::Type{LaplaceDense}) = Cint(0); class_code(::Type{Laplace})   = Cint(1)
class_code(::Type{ConvDiff})     = Cint(2); class_code(::Type{EigValMat}) = Cint(3)
class_code(::Type{RandSPD})      = Cint(4); class_code(::Type{<:MatrixGallery}) = Cint(5)
variant_code(::Type{<:TensorLanczos})       = Cint(0)
variant_code(::Type{<:TensorLanczosReorth}) = Cint(1)
variant_code(::Type{<:TensorArnoldi})       = Cint(2)

const tables_loaded = Ref(false)
"The library's own copy of coefficients_data/ (only needed for `spectral = :library`)."
function load_tables(path::AbstractString = get(ENV, "TENSORKRYLOV_B200_TABLES",
                                                joinpath(dirname(dirname(pathof(TensorKrylov))), "coefficients_data")))
    tables_loaded[] && return nothing
    check(ccall((:tk_tables_load, libtk), Cint, (Cstring,), path))
    tables_loaded[] = true
    return nothing
end

lasterror() = unsafe_string(ccall((:tk_last_error, libtk), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : error("libtensorkrylov_b200: ", lasterror())

function set_operator!(h::Ptr{Cvoid}, s::Int, A::SparseMatrixCSC{Float64, Int64})
    # Julia's SparseMatrixCSC goes over verbatim: 1-based Int64 colptr/rowval
    check(ccall((:tk_set_operator_csc, libtk), Cint,
                (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                h, s - 1, size(A, 1), A.colptr, A.rowval, A.nzval))
end
set_operator!(h::Ptr{Cvoid}, s::Int, A::Matrix{Float64}, uplo::Char = 'F') =
    check(ccall((:tk_set_operator_dense, libtk), Cint, (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Cchar),
                h, s - 1, size(A, 1), A, Cchar(uplo)))     # Cchar(...): ccall does not convert a Char by itself
# Symmetric(R'R, :L) (tensor_struct.jl:77): the library reads the lower triangle of the parent, like the wrapper does
set_operator!(h::Ptr{Cvoid}, s::Int, A::Symmetric{Float64, Matrix{Float64}}) =
    A.uplo == 'L' ? set_operator!(h, s, parent(A), 'L') : set_operator!(h, s, Matrix(A), 'F')
# anything else the gallery or a caller can produce (SymTridiagonal, Diagonal, other index types, views)
set_operator!(h::Ptr{Cvoid}, s::Int, A::AbstractSparseMatrix) = set_operator!(h, s, SparseMatrixCSC{Float64, Int64}(A))
set_operator!(h::Ptr{Cvoid}, s::Int, A::AbstractMatrix)       = set_operator!(h, s, Matrix{Float64}(A), 'F')

"""
    tensorkrylov_b200!(convergence_data, A, b, tol, nmax, orthonormalization_type; device = 0, flags = REFERENCE_H1)

Same contract as `tensorkrylov!`: returns the `KruskalTensor` x on convergence, `nothing` otherwise; on
`CompressedNormBreakdown` prints the reference's message, sets `niterations = k - 1` and `resize!`s the histories.
"""
function tensorkrylov_b200!(convergence_data::ConvergenceData{T}, A::KronMat{matT, U}, b::KronProd{T}, tol::T, nmax::Int,
                            orthonormalization_type::Type{<:TensorDecomposition};
                            device::Int = 0, flags::Cint = TK_FLAG_REFERENCE_H1,
                            spectral::Symbol = :julia) where {matT, T<:Float64, U}
    d  = length(A)
    ns = Int64.(dimensions(A))
    href = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:tk_create, libtk), Cint,
                (Ref{Ptr{Cvoid}}, Int32, Ptr{Int64}, Int32, Int32, Int32, Int32, Int32, Int32, Int32, Int32, Ptr{Cvoid}),
                href, d, ns, nmax, instance_code(U), class_code(A.matrixclass), variant_code(orthonormalization_type),
                flags, device, 0, 1, C_NULL))
    h = href[]
    try
        # operators: the reference aliases one matrix object d times (tensor_struct.jl:208-210)
        first_of = IdDict{Any, Int}()
        if all(A[s] === A[1] for s in 2:d)
            set_operator!(h, 1, A[1])
            check(ccall((:tk_share_operator_all, libtk), Cint, (Ptr{Cvoid}, Int32), h, 0))
        else
        for s in 1:d
            if haskey(first_of, A[s])
                check(ccall((:tk_share_operator, libtk), Cint, (Ptr{Cvoid}, Int32, Int32), h, s - 1, first_of[A[s]] - 1))
            else
                set_operator!(h, s, A[s]); first_of[A[s]] = s
            end
        end
        end
        for s in 1:d
            bs = convert(Vector{Float64}, b[s])   # KronProd{T} = Vector{<:AbstractVector{T}}: views are allowed
            check(ccall((:tk_set_rhs, libtk), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), h, s - 1, bs, length(bs)))
        end
        if spectral === :library
            # the whole schedule inside the library: table lookup and the eigen-extremes of the minors of A_1
            # (tk_schedule; eigenvalues.jl:335-350 + approximation.jl:160-175), no Julia arithmetic at all
            load_tables()
            check(ccall((:tk_schedule, libtk), Cint, (Ptr{Cvoid}, Float64), h, tol))
        else
            # exp-sum schedule: exactly the reference's two update_data! calls, hoisted out of the loop
            # (they depend on A_1, d, tol and k only -- tensor_krylov_method.jl:72-73); lambda_min is then
            # bit-identical to the reference's
            spectraldata = SpectralData{matT, T, U}(A, nmax)
            approxdata   = ApproximationData{T, U}(tol)
            for k in 2:nmax
                update_data!(spectraldata, d, A.matrixclass())
                update_data!(approxdata, spectraldata)
                check(ccall((:tk_set_schedule, libtk), Cint, (Ptr{Cvoid}, Int32, Float64, Int32, Ptr{Float64}, Ptr{Float64}),
                            h, k, spectraldata.λ_min[k], length(approxdata.ω), approxdata.α, approxdata.ω))
            end
        end
        status = Ref{Int32}(0); niter = Ref{Int64}(0); termk = Ref{Int32}(0)
        check(ccall((:tk_solve, libtk), Cint,
                    (Ptr{Cvoid}, Float64, Ref{Int32}, Ref{Int64}, Ref{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    h, tol, status, niter, termk, convergence_data.relative_residual_norm,
                    convergence_data.projected_residual_norm, convergence_data.orthogonality_data))
        if status[] == 2                      # CompressedNormBreakdown, tensor_krylov_method.jl:85-96
            println("Early termination at k = " * string(termk[]) * " due to compressed norm breakdown")
            convergence_data.niterations = niter[]
            resize!(convergence_data, convergence_data.niterations)
            return nothing
        elseif status[] == 0                  # converged, :108-118
            tref = Ref{Int32}(0)
            check(ccall((:tk_solution_rank, libtk), Cint, (Ptr{Cvoid}, Ref{Int32}), h, tref))
            t = Int(tref[])
            x = KruskalTensor{T}(ones(t), [zeros(Int(ns[s]), t) for s in 1:d])
            for s in 1:d
                check(ccall((:tk_get_solution, libtk), Cint,
                            (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Int32),
                            h, s - 1, x.lambda, length(x.lambda), x.fmat[s], length(x.fmat[s]), 0))
            end
            println("Convergence")
            return x
        elseif status[] == 3
            error("NaN in the residual estimate at k = ", termk[])
        end
        println("No convergence")             # :122
        return nothing
    finally
        ccall((:tk_destroy, libtk), Cvoid, (Ptr{Cvoid},), h)
    end
end

"Route the reference's entry point (system.jl:65-83) through the GPU library."
function solve_tensorized_system_b200(system, nmax::Int, orth::Type{<:TensorDecomposition}, tol = 1e-9; kw...)
    convergencedata = ConvergenceData{typeof(tol)}(nmax)
    tensorkrylov_b200!(convergencedata, system.A, system.b, tol, nmax, orth; kw...)
    return convergencedata
end

"""
    TensorKrylovB200.install!()

Make the GPU library THE solve path of the reference package: adds a method of `TensorKrylov.tensorkrylov!` for
`Float64` data.  It is more specific than the package's generic method (tensor_krylov_method.jl:36-43), so
`solve_tensorized_system`, the experiment drivers and the test-suite reach the library without any edit, while
other element types keep falling through to the Julia implementation.
"""
function install!()
    @eval TensorKrylov function tensorkrylov!(
            convergence_data::ConvergenceData{Float64}, A::KronMat{matT, U}, b::KronProd{Float64}, tol::Float64,
            nmax::Int, orthonormalization_type::Type{<:TensorDecomposition},
            mode::Type{<:Mode} = SilentMode) where {matT, U<:Instance}
        return $(tensorkrylov_b200!)(convergence_data, A, b, tol, nmax, orthonormalization_type)
    end
    return nothing
end

export tensorkrylov_b200!, solve_tensorized_system_b200

end # module
