# Reference-side binding of libgraphbix_cuda.so (include/graphbix_cuda.h): the two functions a maintainer of graphBIX
# includes from sbm.jl / mod.jl in place of the bodies of `EM(...)`.  Julia is not available in the image this library
# is built and tested in, so this file is a line-for-line translation of the ctypes binding the test-suite exercises
# (graphbix_b200/_lib.py, graphbix_b200/engine.py: EM, EM_mod) and has not been run.  Written for Julia >= 1.0
# (the reference targets 0.6: replace `Ref{...}` struct passing by `pointer_from_objref` there).
#
#   sbm.jl:918   (gr,Crs,PSI,FE,CVBayes,CVGP,CVGT,CVMAP,varCVBayes,varCVGP,varCVGT,varCVMAP,fail,cnv,itrnum) =
#                    EM_cuda(Ntot,B,links,maxCrs,gr0,Crs0,degreecorrection,itrmax,priorlearning,learningrate)
#   mod.jl:817   (excm,alpha,beta,omegain,omegaout,PSI,FE,CVBayes,CVGP,CVGT,CVMAP,varCVBayes,varCVGP,varCVGT,varCVMAP,cnv,itrnum) =
#                    EM_mod_cuda(gr,Ntot,B,links,maxomega,itrmax,betafrac,learning,priorlearning,dc,learningrate)
# integration/sbm.jl.patch and integration/mod.jl.patch are the diffs against the reference's src/ that make exactly
# these two replacements (`patch -p1` in the reference's root; copy this file next to sbm.jl / mod.jl).

const libgbx = "libgraphbix_cuda"

mutable struct GbxSbmOptions               # struct gbx_sbm_options
    degreecorrection::Cint
    priorlearning::Cint
    learningrate::Cdouble
    maxCrs::Cdouble
    itrmax::Cint
    bp_conv_threshold::Cdouble
    damping::Cdouble
    check_every::Cint
    mstep_exact::Cint
    GbxSbmOptions() = new()
end

mutable struct GbxSbmResult                # struct gbx_sbm_result
    FE::Cdouble; CVBayes::Cdouble; CVGP::Cdouble; CVGT::Cdouble; CVMAP::Cdouble
    varCVBayes::Cdouble; varCVGP::Cdouble; varCVGT::Cdouble; varCVMAP::Cdouble
    fail::Cint; cnv::Cint; itrnum::Cint; itrs_done::Cint; nonfinite::Cint
    BPconv::Cdouble; max_dpsi::Cdouble
    GbxSbmResult() = new()
end

mutable struct GbxModOptions               # struct gbx_mod_options
    dc::Cint
    learning::Cint
    priorlearning::Cint
    learningrate::Cdouble
    maxomega::Cdouble
    itrmax::Cint
    betafrac::Cdouble
    bp_conv_threshold::Cdouble
    damping::Cdouble
    check_every::Cint
    GbxModOptions() = new()
end

mutable struct GbxModResult                # struct gbx_mod_result
    excm::Cdouble; alpha::Cdouble; beta::Cdouble; omegain::Cdouble; omegaout::Cdouble
    FE::Cdouble; CVBayes::Cdouble; CVGP::Cdouble; CVGT::Cdouble; CVMAP::Cdouble
    varCVBayes::Cdouble; varCVGP::Cdouble; varCVGT::Cdouble; varCVMAP::Cdouble
    cnv::Cint; itrnum::Cint; itrs_done::Cint; nonfinite::Cint
    BPconv::Cdouble; max_dpsi::Cdouble
    GbxModResult() = new()
end

gbxcheck(rc) = rc == 0 || error("libgraphbix_cuda: " * unsafe_string(ccall((:gbx_last_error, libgbx), Cstring, ())))

# sbm.jl: the hyper-parameter initial conditions (sbm.jl:279-302, including Laplacian + Kmeans) stay in Julia and hand
# over (gr0, Crs0); the random initial messages (sbm.jl:319-323) are drawn by the library from `seed`.
function EM_cuda(Ntot, B, links::Matrix{Int64}, maxCrs, gr0, Crs0, degreecorrection, itrmax, priorlearning,
                 learningrate; device = 0, seed = rand(UInt64))
    Ktot = size(links, 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    # `links` goes over as Julia stores it: K x 2, column-major, 1-based, both directions of every edge
    gbxcheck(ccall((:gbx_sbm_create, libgbx), Cint, (Ref{Ptr{Cvoid}}, Cint, Int64, Int64, Ptr{Int64}, Cint),
                   h, device, Ntot, Ktot, links, B))
    try
        opt = GbxSbmOptions()
        ccall((:gbx_sbm_default_options, libgbx), Cvoid, (Ref{GbxSbmOptions},), opt)
        opt.degreecorrection = degreecorrection; opt.priorlearning = priorlearning
        opt.learningrate = learningrate; opt.maxCrs = maxCrs; opt.itrmax = itrmax
        gbxcheck(ccall((:gbx_sbm_set_options, libgbx), Cint, (Ptr{Cvoid}, Ref{GbxSbmOptions}), h[], opt))
        Crs_rm = permutedims(Matrix{Float64}(Crs0))                  # the C ABI is row-major (Crs is symmetric anyway)
        gbxcheck(ccall((:gbx_sbm_init, libgbx), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, UInt64),
                       h[], vec(Float64.(gr0)), Crs_rm, C_NULL, seed))
        res = GbxSbmResult()
        gbxcheck(ccall((:gbx_sbm_em, libgbx), Cint, (Ptr{Cvoid}, Ref{GbxSbmResult}), h[], res))
        gr = zeros(B); Crs = zeros(B, B); PSIt = zeros(B, Ntot)      # PSI comes back as double[N][q]
        gbxcheck(ccall((:gbx_sbm_get_state, libgbx), Cint,
                       (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                       h[], gr, Crs, PSIt, C_NULL, C_NULL))
        return (gr', permutedims(Crs), permutedims(PSIt), res.FE, res.CVBayes, res.CVGP, res.CVGT, res.CVMAP,
                res.varCVBayes, res.varCVGP, res.varCVGT, res.varCVMAP, res.fail != 0, res.cnv != 0, Int(res.itrnum))
    finally
        ccall((:gbx_sbm_destroy, libgbx), Cint, (Ptr{Cvoid},), h[])
    end
end

# mod.jl: same graph conventions; `gr` is the prior the driver passes in (mod.jl:807: ones(1,B)/B).
function EM_mod_cuda(gr, Ntot, B, links::Matrix{Int64}, maxomega, itrmax, betafrac, learning, priorlearning, dc,
                     learningrate; device = 0, seed = rand(UInt64))
    Ktot = size(links, 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    gbxcheck(ccall((:gbx_mod_create, libgbx), Cint, (Ref{Ptr{Cvoid}}, Cint, Int64, Int64, Ptr{Int64}, Cint),
                   h, device, Ntot, Ktot, links, B))
    try
        opt = GbxModOptions()
        ccall((:gbx_mod_default_options, libgbx), Cvoid, (Ref{GbxModOptions},), opt)
        opt.dc = dc; opt.learning = learning; opt.priorlearning = priorlearning; opt.learningrate = learningrate
        opt.maxomega = maxomega; opt.itrmax = itrmax; opt.betafrac = betafrac
        gbxcheck(ccall((:gbx_mod_set_options, libgbx), Cint, (Ptr{Cvoid}, Ref{GbxModOptions}), h[], opt))
        gbxcheck(ccall((:gbx_mod_init, libgbx), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, UInt64),
                       h[], vec(Float64.(gr)), C_NULL, seed))
        res = GbxModResult()
        gbxcheck(ccall((:gbx_mod_em, libgbx), Cint, (Ptr{Cvoid}, Ref{GbxModResult}), h[], res))
        PSIt = zeros(B, Ntot)
        gbxcheck(ccall((:gbx_mod_get_state, libgbx), Cint,
                       (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                       h[], C_NULL, PSIt, C_NULL, C_NULL))
        return (res.excm, res.alpha, res.beta, res.omegain, res.omegaout, permutedims(PSIt), res.FE, res.CVBayes,
                res.CVGP, res.CVGT, res.CVMAP, res.varCVBayes, res.varCVGP, res.varCVGT, res.varCVMAP, res.cnv != 0,
                Int(res.itrnum))
    finally
        ccall((:gbx_mod_destroy, libgbx), Cint, (Ptr{Cvoid},), h[])
    end
end
