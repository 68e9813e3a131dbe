# runtests_b200.jl -- Julia-side tests of the libtensorkrylov_b200.so drop-in.
#
# Run from a Julia environment that has TensorKrylov.jl (the reference) and Kronecker.jl, on a machine with a B200:
#
#     TENSORKRYLOV_B200_LIB=/path/to/libtensorkrylov_b200.so \
#     TENSORKRYLOV_B200_TABLES=/path/to/TensorKrylov.jl/coefficients_data \
#     julia --project=/path/to/TensorKrylov.jl runtests_b200.jl
#
# What is ported from the reference's own suite:
#   * test/tensor_krylov_method.jl:31-45   "Symmetric example"  (d=5, n=200 Laplace, TensorLanczosReorth, nmax=199): the
#     reference only runs and displays it; here the same system is solved by the stock Julia `tensorkrylov!` AND by the
#     library, and the two ConvergenceData are compared (tolerance model of SURVEY.md 8c).
#   * test/tensor_krylov_method.jl:47-61   "Nonsymmetric example" (ConvDiff / TensorArnoldi; the solve is commented out
#     in the reference): same comparison.
#   * test/utils.jl:88-185                 the dense Kronecker checks of MVnorm / tensorinnerprod / compressed_residual
#     (d=3, n=15): the library's estimator terms for ITS compressed solution against explicit `kroneckersum` / `kron`
#     arithmetic (Kronecker.jl), and the true residual of the returned KruskalTensor.
#   * `install!()`: after it, the package's own `solve_tensorized_system` reaches the library unchanged.
#
# This file cannot be executed in the build image (no Julia there).  The same checks run there from Python
# (tests/test_gpu_parity.py::test_true_residual_of_returned_solution, test_compress_and_residual_phases, the corpus).
using Test, Random, LinearAlgebra, SparseArrays
using Kronecker
using TensorKrylov
using TensorKrylov: KronMat, KronProd, ConvergenceData, KruskalTensor, TensorizedSystem, TensorLanczosReorth,
                    TensorLanczos, TensorArnoldi, SymInstance, NonSymInstance, Laplace, ConvDiff, random_rhs,
                    solve_tensorized_system, assemble_matrix, kroneckervectorize

include(joinpath(@__DIR__, "TensorKrylovB200.jl"))
using .TensorKrylovB200
const B = TensorKrylovB200
const libtk = B.libtk

Random.seed!(12345)                      # test/runtests.jl:9

tables = get(ENV, "TENSORKRYLOV_B200_TABLES", joinpath(dirname(dirname(pathof(TensorKrylov))), "coefficients_data"))
B.check(ccall((:tk_tables_load, libtk), Cint, (Cstring,), tables))

# relres^2 = (boundary + r_comp)/||b||^2 with r_comp a cancellation of O(||b||^2) terms, each reproduced to 1e-11
relres_close(a, b; scale = 4.0) = maximum(abs.(a .^ 2 .- b .^ 2)) <= 1e-11 * scale

"Stock Julia solve and library solve of the same system; returns both ConvergenceData."
function both(instance, class, orth, d, n, nmax, tol)
    A  = KronMat{instance}(d, n, class)
    bs = rand(n)
    # TensorizedSystem normalises b in place (system.jl:33-37): give each arm its own copy of the same vector
    sysref = TensorizedSystem{instance}(A, [copy(bs) for _ in 1:1][ones(Int, d)])
    syslib = TensorizedSystem{instance}(A, [copy(bs) for _ in 1:1][ones(Int, d)])
    cdref  = solve_tensorized_system(sysref, nmax, orth, tol)
    cdlib  = B.solve_tensorized_system_b200(syslib, nmax, orth, tol)
    return cdref, cdlib
end

@testset "libtensorkrylov_b200 drop-in" begin

    @testset "Symmetric example (test/tensor_krylov_method.jl:31-45)" begin
        cdref, cdlib = both(SymInstance, Laplace, TensorLanczosReorth, 5, 200, 199, 1e-9)
        @test cdlib.niterations == cdref.niterations
        @test length(cdlib.relative_residual_norm) == length(cdref.relative_residual_norm)
        @test cdlib.relative_residual_norm[1] == 1.0 && cdlib.projected_residual_norm[1] == 1.0   # convergence.jl:11-20
        k = 2:min(60, cdref.niterations)      # beyond k ~ 60 r_comp is a noise-level cancellation in BOTH implementations
        @test relres_close(cdlib.relative_residual_norm[k], cdref.relative_residual_norm[k])
        @test maximum(abs.(cdlib.relative_residual_norm[2:10] .- cdref.relative_residual_norm[2:10]) ./
                      cdref.relative_residual_norm[2:10]) < 1e-10
        @test maximum(abs.(cdlib.orthogonality_data[k] .- cdref.orthogonality_data[k])) < 1e-12
    end

    @testset "Nonsymmetric example (test/tensor_krylov_method.jl:47-61)" begin
        cdref, cdlib = both(NonSymInstance, ConvDiff, TensorArnoldi, 5, 200, 120, 1e-9)
        k = 2:min(100, cdref.niterations, cdlib.niterations)
        @test relres_close(cdlib.relative_residual_norm[k], cdref.relative_residual_norm[k])
        @test maximum(abs.(cdlib.relative_residual_norm[k] .- cdref.relative_residual_norm[k]) ./
                      cdref.relative_residual_norm[k]) < 1e-9
    end

    @testset "Dense Kronecker checks (test/utils.jl:88-185)" begin
        # d = 3, n = 15 as in the reference's tests; distinct b_s, every mode its own H_s (flags = 0)
        d, n, nmax, tol = 3, 15, 10, 1e-9
        h  = inv(n + 1)
        Mi = [sparse(inv(h^2) .* Tridiagonal(-ones(n - 1), 2ones(n), -ones(n - 1))) for _ in 1:d]   # test/utils.jl:172
        b  = [rand(n) for _ in 1:d]
        for s in 1:d
            b[s] .*= inv(norm(b[s]))
        end
        href = Ref{Ptr{Cvoid}}(C_NULL)
        ns   = fill(Int64(n), d)
        B.check(ccall((:tk_create, libtk), Cint,
                      (Ref{Ptr{Cvoid}}, Int32, Ptr{Int64}, Int32, Int32, Int32, Int32, Int32, Int32, Int32, Int32, Ptr{Cvoid}),
                      href, d, ns, nmax, 0, 1, 1, 0, 0, 0, 1, C_NULL))          # Sym, Laplace, LanczosReorth, per-mode H_s
        hd = href[]
        try
            for s in 1:d
                B.set_operator!(hd, s, Mi[s])
                B.check(ccall((:tk_set_rhs, libtk), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), hd, s - 1, b[s], n))
            end
            B.check(ccall((:tk_schedule_laplace, libtk), Cint, (Ptr{Cvoid}, Float64), hd, tol))
            B.check(ccall((:tk_begin, libtk), Cint, (Ptr{Cvoid},), hd))
            out8 = zeros(8)
            for k in 2:nmax
                B.check(ccall((:tk_step_bases, libtk), Cint, (Ptr{Cvoid}, Int32), hd, k))
                B.check(ccall((:tk_compress, libtk), Cint, (Ptr{Cvoid}, Int32), hd, k))
                B.check(ccall((:tk_residual, libtk), Cint, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}), hd, k, 0.0, out8))
                # the library's H_s, b~_s and Y_s of this iteration
                tref = Ref{Int32}(0)
                B.check(ccall((:tk_get_Y, libtk), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ref{Int32}), hd, 0, k, C_NULL, tref))
                t  = Int(tref[])
                Hs = Matrix{Float64}[]; Ys = Matrix{Float64}[]; bts = Vector{Float64}[]
                for s in 1:d
                    Hfull = zeros(nmax + 1, nmax + 1)
                    B.check(ccall((:tk_get_H, libtk), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), hd, s - 1, Hfull))
                    push!(Hs, Hfull[1:k, 1:k])
                    Y = zeros(k, t)
                    B.check(ccall((:tk_get_Y, libtk), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ref{Int32}), hd, s - 1, k, Y, tref))
                    push!(Ys, Y)
                    bt = zeros(nmax + 1)
                    B.check(ccall((:tk_get_bt, libtk), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), hd, s - 1, bt))
                    push!(bts, bt[1:k])
                end
                # y.lambda = omega / lambda_min (tensor_krylov_method.jl:23): recover it from the tables like the library
                lmin = Ref{Float64}(0.0); lmax = Ref{Float64}(0.0)
                B.check(ccall((:tk_laplace_extremes, libtk), Cint, (Int32, Int64, Int32, Ref{Float64}, Ref{Float64}), d, n, k, lmin, lmax))
                om = zeros(64); al = zeros(64); tt = Ref{Int32}(0); dg = Ref{Int32}(0); od = Ref{Int32}(0)
                B.check(ccall((:tk_tables_sym_lookup, libtk), Cint,
                              (Float64, Float64, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ptr{Float64}, Ptr{Float64}),
                              lmax[] * inv(lmin[]), tol, tt, dg, od, om, al))
                @test Int(tt[]) == t
                lam = om[1:t] .* inv(lmin[])
                y   = KruskalTensor{Float64}(lam, Ys)
                yv  = kroneckervectorize(y)                             # tensor_struct.jl:361-384
                Hk  = kroneckersum(reverse(Hs)...)                      # test_utils.jl:66; kroneckervectorize has mode 1 fastest,
                                                                        # kroneckersum(A, B, ..) puts its FIRST argument slowest
                btv = kron(reverse(bts)...)
                Hy  = Hk * yv
                @test isapprox(out8[1], dot(Hy, Hy); rtol = 1e-11)      # MVnorm                 (test/utils.jl:112)
                @test isapprox(out8[2], dot(Hy, btv); rtol = 1e-11)     # tensorinnerprod        (test/utils.jl:123)
                @test isapprox(out8[3], dot(btv, btv); rtol = 1e-11)    # kronproddot
                scale = abs(out8[1]) + 2abs(out8[2]) + abs(out8[3])
                @test abs(out8[5] - norm(Hy - btv)^2) <= 1e-11 * scale  # compressed_residual    (test/utils.jl:126)
            end
            # the true residual of the returned iterate equals the estimator's value (Lemma 3.4 is an identity for y)
            tref = Ref{Int32}(0)
            B.check(ccall((:tk_solution_rank, libtk), Cint, (Ptr{Cvoid}, Ref{Int32}), hd, tref))
            t = Int(tref[])
            x = KruskalTensor{Float64}(ones(t), [zeros(n, t) for _ in 1:d])
            for s in 1:d
                B.check(ccall((:tk_get_solution, libtk), Cint,
                              (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Int32),
                              hd, s - 1, x.lambda, t, x.fmat[s], n * t, 1))
            end
            Ad   = kroneckersum(reverse(Mi)...)
            bd   = kron(reverse(b)...)
            true_rel = norm(Ad * kroneckervectorize(x) - bd) / norm(bd)
            @test isapprox(out8[6] / norm(bd), true_rel; rtol = 1e-6)
        finally
            ccall((:tk_destroy, libtk), Cvoid, (Ptr{Cvoid},), hd)
        end
    end

    @testset "install!() routes the package's own entry point" begin
        B.install!()
        d, n, nmax = 5, 200, 40
        A   = KronMat{SymInstance}(d, n, Laplace)
        bs  = rand(n)
        sys = TensorizedSystem{SymInstance}(A, [copy(bs)][ones(Int, d)])
        cd  = solve_tensorized_system(sys, nmax, TensorLanczosReorth, 1e-9)        # now the library (Float64 method)
        sys2 = TensorizedSystem{SymInstance}(A, [copy(bs)][ones(Int, d)])
        cd2  = B.solve_tensorized_system_b200(sys2, nmax, TensorLanczosReorth, 1e-9)
        @test cd.relative_residual_norm == cd2.relative_residual_norm               # same library, same inputs: bit-identical
        @test cd.niterations == cd2.niterations
    end

end
