# Reference-side binding of libgraphbix_cuda.so (include/graphbix_cuda.h): the functions a maintainer of graphBIX
# includes from sbm.jl / mod.jl (and the labeled-SBM notebook, at the end of this file) in place of the bodies of `EM(...)`.  Julia is not available in the image this library
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

# ---- labeled SBM of the notebook (src/labeled_sbm/labeled-sbmBIX_0.2.3.8_Julia0.6.2.ipynb, cell 0) -----------------
#   notebook MAIN:  (gr,affinity,PSI,FE,CVBayes,CVGP,CVGT,CVMAP,varCVBayes,varCVGP,varCVGT,varCVMAP,cnv,itrnum) =
#                       EM_labeled_cuda(Ntot,Larray,B,G,links,itrmax,BPconvthreshold,learning,priorlearning,dc,
#                                       learningrate,REDfrac,assortativity,Omega)
# in place of EM(...) with the same arguments.  `links` is the notebook's K x 3 Int64 matrix (i, j, label) after
# labeledsimplegraph; `affinity` comes back as the notebook's Dict(alpha => B x B matrix).  Like the two stubs above
# this is a translation of the tested ctypes binding (graphbix_b200/labeled.py: EM_labeled, EM_labeled_batch).

mutable struct GbxLsbmOptions              # struct gbx_lsbm_options
    dc::Cint
    learning::Cint
    priorlearning::Cint
    learningrate::Cdouble
    itrmax::Cint
    bp_conv_threshold::Cdouble
    REDfrac::Cdouble
    damping::Cdouble
    GbxLsbmOptions() = new()
end

struct GbxLsbmResult                       # struct gbx_lsbm_result (immutable: it also travels in Vectors, one per graph)
    FE::Cdouble; CVBayes::Cdouble; CVGP::Cdouble; CVGT::Cdouble; CVMAP::Cdouble
    varCVBayes::Cdouble; varCVGP::Cdouble; varCVGT::Cdouble; varCVMAP::Cdouble
    cnv::Cint; itrnum::Cint; itrs_done::Cint; nonfinite::Cint
    BPconv::Cdouble
end

# The notebook's "initial state" block of EM() (affinity[k] from Larray, assortativity, Omega; the last label flat when
# REDfrac != 0) as B x B x G, which is the C ABI's double[G][B][B] (the matrices are symmetric).
function initial_affinity_cuda(Ntot, Ktot, Larray, B, G, dc, REDfrac, assortativity, Omega)
    cm = 2 .* Larray ./ Ntot
    nrm = dc ? Ktot : Ntot
    aff = zeros(B, B, G)
    for k = 1:G
        if REDfrac != 0 && k == G
            aff[:, :, k] .= cm[k] / nrm
            continue
        end
        delta = assortativity[k] == "a" ? 0.99 * cm[k] / Omega :
                assortativity[k] == "d" ? -0.99 * cm[k] / (1 - Omega) : 0.0
        aff[:, :, k] .= (cm[k] - delta * Omega) / nrm
        for r = 1:B
            aff[r, r, k] += delta / nrm
        end
    end
    return aff
end

function lsbm_set_options(h, itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate, REDfrac)
    opt = GbxLsbmOptions()
    ccall((:gbx_lsbm_default_options, libgbx), Cvoid, (Ref{GbxLsbmOptions},), opt)
    opt.dc = dc; opt.learning = learning; opt.priorlearning = priorlearning; opt.learningrate = learningrate
    opt.itrmax = itrmax; opt.bp_conv_threshold = BPconvthreshold; opt.REDfrac = REDfrac
    gbxcheck(ccall((:gbx_lsbm_set_options, libgbx), Cint, (Ptr{Cvoid}, Ref{GbxLsbmOptions}), h, opt))
end

function lsbm_result_tuple(h, g, Ntot, B, G, res::GbxLsbmResult)
    gr = zeros(B); aff = zeros(B, B, G); PSIt = zeros(B, Ntot)       # PSI comes back as double[N][q]
    gbxcheck(ccall((:gbx_lsbm_get_state_of, libgbx), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                   h, g, gr, aff, PSIt, C_NULL))
    affinity = Dict(alpha => permutedims(aff[:, :, alpha]) for alpha = 1:G)
    return (gr', affinity, permutedims(PSIt), res.FE, res.CVBayes, res.CVGP, res.CVGT, res.CVMAP, res.varCVBayes,
            res.varCVGP, res.varCVGT, res.varCVMAP, res.cnv != 0, Int(res.itrnum))
end

function EM_labeled_cuda(Ntot, Larray, B, G, links::Matrix{Int64}, itrmax, BPconvthreshold, learning, priorlearning, dc,
                         learningrate, REDfrac, assortativity, Omega; device = 0, seed = rand(UInt64))
    Ktot = size(links, 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    gbxcheck(ccall((:gbx_lsbm_create, libgbx), Cint, (Ref{Ptr{Cvoid}}, Cint, Int64, Int64, Ptr{Int64}, Cint, Cint),
                   h, device, Ntot, Ktot, links, B, G))
    try
        lsbm_set_options(h[], itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate, REDfrac)
        aff0 = initial_affinity_cuda(Ntot, Ktot, Larray, B, G, dc, REDfrac, assortativity, Omega)
        # the rand(1,B) messages of the notebook are drawn on the device: row k of the graph from (seed, k)
        gbxcheck(ccall((:gbx_lsbm_init_random, libgbx), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt64}),
                       h[], fill(1.0 / B, B), aff0, UInt64[seed]))
        res = Vector{GbxLsbmResult}(undef, 1)
        gbxcheck(ccall((:gbx_lsbm_em, libgbx), Cint, (Ptr{Cvoid}, Ptr{GbxLsbmResult}), h[], res))
        return lsbm_result_tuple(h[], 0, Ntot, B, G, res[1])
    finally
        ccall((:gbx_lsbm_destroy, libgbx), Cint, (Ptr{Cvoid},), h[])
    end
end

# Many graphs at once (a detectability sweep: the notebook's MAIN loops EM(...) over samples): ONE handle, every kernel
# launch serves all graphs, each graph stops at its own iteration.  graphs[g] = (Ntot, Larray, links); equal B and G.
# Returns one 14-tuple per graph, each what EM_labeled_cuda returns for that graph from the same seed.
function EM_labeled_cuda_batch(graphs, B, G, itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate, REDfrac,
                               assortativity, Omega; device = 0, seeds = rand(UInt64, length(graphs)))
    count = length(graphs)
    ns = Int64[g[1] for g in graphs]
    Ks = Int64[size(g[3], 1) for g in graphs]
    linkmats = Matrix{Int64}[g[3] for g in graphs]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve linkmats begin
        ptrs = Ptr{Int64}[pointer(l) for l in linkmats]
        gbxcheck(ccall((:gbx_lsbm_create_batch, libgbx), Cint,
                       (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Ptr{Int64}}, Cint, Cint),
                       h, device, count, ns, Ks, ptrs, B, G))
    end
    try
        lsbm_set_options(h[], itrmax, BPconvthreshold, learning, priorlearning, dc, learningrate, REDfrac)
        aff0 = zeros(B, B, G, count)                                 # double[count][G][B][B]
        for g = 1:count
            aff0[:, :, :, g] = initial_affinity_cuda(ns[g], Ks[g], graphs[g][2], B, G, dc, REDfrac, assortativity, Omega)
        end
        gbxcheck(ccall((:gbx_lsbm_init_random, libgbx), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt64}),
                       h[], fill(1.0 / B, B, count), aff0, Vector{UInt64}(seeds)))
        res = Vector{GbxLsbmResult}(undef, count)
        gbxcheck(ccall((:gbx_lsbm_em, libgbx), Cint, (Ptr{Cvoid}, Ptr{GbxLsbmResult}), h[], res))
        return [lsbm_result_tuple(h[], g - 1, ns[g], B, G, res[g]) for g = 1:count]
    finally
        ccall((:gbx_lsbm_destroy, libgbx), Cint, (Ptr{Cvoid},), h[])
    end
end
