# gpu_separator.jl -- the reference-side binding of the B200 separation library (libktn.so, include/ktn.h).
#
# A Katana.jl maintainer adds this file to src/ and `include`s it from src/Katana.jl after separators.jl and nlpeval.jl.  It is the
# code of INTEGRATION.md sections 3, 4 and 6 verbatim (tests/test_abi.py keeps the two in step and checks every ccall name and
# argument count against the header and the exported symbols; tests/test_julia_shim_rules.py runs the shim's flattening rules,
# ported 1:1 to Python, through the C ABI on the reference's test problems).  NOT executed in this repository: the build image
# has no julia binary.
#
mutable struct KatanaGPUSeparator <: AbstractKatanaSeparator
    handle :: Ptr{Cvoid}
    num_var :: Int
    num_constr :: Int
    ngpus :: Int                    # > 1: the ONE separator shards its rows over that many devices of this process
    linear_model                    # JuMP model the cuts' variables belong to (src/separators.jl:62,86)
    cols :: Vector{Vector{Int}}     # per row: Jacobian columns, 1-based (the reference's sp_cols, src/separators.jl:92-100)
    xstar :: Vector{Float64}
    g :: Vector{Float64}            # filled lazily for the per-row isconstrsat hook
    have_g :: Bool
    KatanaGPUSeparator(; ngpus = 1) = new(C_NULL, 0, 0, ngpus, nothing, Vector{Int}[], Float64[], Float64[], false)
end

struct KtnOptions            # include/ktn.h: ktn_options
    struct_size::Int32; device::Int32; f_tol::Float64; cut_coef_rng::Float64; topk::Int64; flags::Int32; reserved::Int32
    ngpus::Int32; devices::NTuple{16, Int32}
end

struct KtnCutView            # include/ktn.h: ktn_cut_view
    n_cuts::Int64; nnz::Int64
    row_id::Ptr{Int64}; row_ptr::Ptr{Int64}; col::Ptr{Int32}; val::Ptr{Float64}
    lo::Ptr{Float64}; hi::Ptr{Float64}; g::Ptr{Float64}; viol::Ptr{Float64}; bconst::Ptr{Float64}
end

check(sep, rc, what) = rc < 0 && error("$what failed ($rc): " *
    unsafe_string(ccall((:ktn_last_error, libktn), Cstring, (Ptr{Cvoid},), sep.handle)))

# op codes of the wire format (include/ktn.h KTN_OP_*)
const KTN_OPS = Dict(:+ => 2, :- => 3, :* => 4, :/ => 5, :^ => 6, :exp => 8, :log => 9, :sqrt => 10, :abs => 11, :sin => 12, :cos => 13,
                     :ifelse => 14, :<= => 15, :< => 16, :>= => 17, :> => 18, :(==) => 19)

# Flatten one expression tree over x[i] (MathProgBase constr_expr / obj_expr) into prefix arrays.
function flatten!(op::Vector{Int32}, arg::Vector{Int32}, val::Vector{Float64}, ex)
    if ex isa Number
        push!(op, 0); push!(arg, 0); push!(val, Float64(ex))
    elseif ex isa Expr && ex.head == :ref                     # x[i]
        push!(op, 1); push!(arg, Int32(ex.args[2] - 1)); push!(val, 0.0)
    elseif ex isa Expr && ex.head == :comparison && length(ex.args) == 3     # a <= b inside ifelse (Julia 0.5 / 0.6 parse it as :comparison)
        push!(op, KTN_OPS[ex.args[2]]); push!(arg, 2); push!(val, 0.0)
        flatten!(op, arg, val, ex.args[1]); flatten!(op, arg, val, ex.args[3])
    elseif ex isa Expr && ex.head == :call
        f, a = ex.args[1], ex.args[2:end]
        if f == :- && length(a) == 1                           # unary minus
            push!(op, 7); push!(arg, 1); push!(val, 0.0)
        else
            push!(op, KTN_OPS[f]); push!(arg, Int32(length(a))); push!(val, 0.0)
        end
        foreach(c -> flatten!(op, arg, val, c), a)
    else
        error("unsupported expression node $ex")
    end
end

# The body g_i(x) of a MathProgBase constraint expression: `body <= rhs`, `body >= rhs`, `body == rhs` are :call nodes with the
# body first; a two-sided row `lb <= body <= ub` is a :comparison node with the body in the middle.  Bounds travel separately.
constr_body(c::Expr) = c.head == :comparison ? c.args[3] : c.args[2]

# The epigraph wrapper hands out expression graphs too (src/nlpeval.jl:23 advertises [:Grad, :Jac] only): rows 1 .. num_constr - 1
# are the wrapped evaluator's, the last row is f(x[1:n]) - x[n+1] (src/nlpeval.jl:35,42-45) and is never linear.
MathProgBase.features_available(d::EpigraphNLPEvaluator) = [:Grad, :Jac, :ExprGraph]
MathProgBase.constr_expr(d::EpigraphNLPEvaluator, i) = i < d.num_constr ? MathProgBase.constr_expr(d.nlpeval, i) :
    Expr(:call, :<=, Expr(:call, :-, MathProgBase.obj_expr(d.nlpeval), Expr(:ref, :x, d.num_var)), 0.0)
MathProgBase.isconstrlinear(d::EpigraphNLPEvaluator, i) = i < d.num_constr && MathProgBase.isconstrlinear(d.nlpeval, i)

# LP / QP / QCQP route (src/solver.jl:46): a model without @NL macros arrives wrapped in MathProgBase's NonlinearToLPQPBridge, whose
# evaluator LPQPEvaluator offers [:Grad, :Jac, :Hess] only.  A and Q are captured directly and handed out as expression graphs of
# the shape JuMP prints for quadratic expressions, +(q*x[i]*x[j] ..., a*x[k] ...), merged and sorted.  (Field names as in MathProgBase
# 0.7 SolverInterface/nonlinear_to_lpqp.jl -- third-party, not vendored in the reference: adjust to the installed version.  The
# reference's Jacobian on this route has two entries per quadratic term in COO order; here a cut's columns are the merged row.)
const LPQPEvaluator = MathProgBase.SolverInterface.LPQPEvaluator
MathProgBase.features_available(d::LPQPEvaluator) = [:Grad, :Jac, :Hess, :ExprGraph]
MathProgBase.initialize(d::LPQPEvaluator, feats::Vector{Symbol}) =
    all(f -> f in MathProgBase.features_available(d), feats) || error("Unsupported feature in $feats")
function lpqp_expr(lin::Dict{Int,Float64}, quad::Dict{Tuple{Int,Int},Float64})
    ex = Expr(:call, :+)
    for k in sort(collect(keys(quad))); push!(ex.args, Expr(:call, :*, quad[k], Expr(:ref, :x, k[1]), Expr(:ref, :x, k[2]))); end
    for k in sort(collect(keys(lin))); push!(ex.args, Expr(:call, :*, lin[k], Expr(:ref, :x, k))); end
    length(ex.args) == 1 && push!(ex.args, 0.0)
    ex
end
function MathProgBase.constr_expr(d::LPQPEvaluator, i)
    lin, quad = Dict{Int,Float64}(), Dict{Tuple{Int,Int},Float64}()
    nlin = size(d.A, 1)
    if i <= nlin                                               # row i of A
        rows, vals = rowvals(d.A), nonzeros(d.A)
        for j in 1:size(d.A, 2), p in nzrange(d.A, j)
            rows[p] == i && (lin[j] = get(lin, j, 0.0) + vals[p])
        end
    else                                                       # addquadconstr!: sum linearval * x + sum quadval * x[row] * x[col], entries as given
        q = d.Qconstr[i - nlin]
        for (j, v) in zip(q.linearidx, q.linearval); lin[j] = get(lin, j, 0.0) + v; end
        for (a, b, v) in zip(q.quadrowidx, q.quadcolidx, q.quadval); k = (min(a, b), max(a, b)); quad[k] = get(quad, k, 0.0) + v; end
    end
    Expr(:call, :<=, lpqp_expr(lin, quad), 0.0)                # only the body is read (constr_body); bounds travel separately
end
function MathProgBase.obj_expr(d::LPQPEvaluator)               # c'x + 0.5 x'Qx, Q given by one triangle (setquadobj!)
    lin = Dict{Int,Float64}(j => v for (j, v) in enumerate(d.c) if v != 0.0)
    quad = Dict{Tuple{Int,Int},Float64}()
    for (a, b, v) in zip(d.Qi, d.Qj, d.Qv); k = (min(a, b), max(a, b)); quad[k] = get(quad, k, 0.0) + (a == b ? 0.5 * v : v); end
    lpqp_expr(lin, quad)
end

# initialize!(sep, linear_model, num_var, num_constr, oracle)           -- src/separators.jl:81-107
function initialize!(sep::KatanaGPUSeparator, linear_model, num_var::Int, num_constr::Int, oracle)
    MathProgBase.initialize(oracle, [:ExprGraph])              # the separator initialises the oracle itself (:88)
    if sep.handle == C_NULL                                    # one separator is reused across models (test/runtests.jl:24)
        # flags: 1 = lean views.  Large models on ONE device run as two pipelined shards of that device (flags |= 4, the device listed
        # twice): the transfer of the first shard's cuts overlaps the second shard's kernels (measured default of the Python twin,
        # katana.jl_b200/separators.py: between 750 000 and 4 000 000 rows; decided by the first model the separator sees)
        n_nl = count(i -> !MathProgBase.isconstrlinear(oracle, i), 1:num_constr)      # the rows a round tests and cuts
        pipe = sep.ngpus <= 1 && 750_000 <= n_nl <= 4_000_000
        devs = ntuple(i -> Int32(pipe && i <= 2 ? 0 : -1), 16)
        o = Ref(KtnOptions(sizeof(KtnOptions), -1, 1e-6, 1e9, 0, pipe ? 5 : 1, 0, pipe ? 2 : (sep.ngpus > 1 ? sep.ngpus : 0), devs))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ktn_create, libktn), Cint, (Ref{KtnOptions}, Ref{Ptr{Cvoid}}), o, h)
        rc == 0 || error("ktn_create failed ($rc): no B200 / CUDA device?")
        sep.handle = h[]
    end
    check(sep, ccall((:ktn_load_begin, libktn), Cint, (Ptr{Cvoid}, Int64, Int64), sep.handle, num_var, num_constr), "ktn_load_begin")
    op, arg, val, eptr = Int32[], Int32[], Float64[], Int64[0]
    flags = zeros(UInt8, num_constr)
    dense_row = oracle isa EpigraphNLPEvaluator ? num_constr : 0          # its Jacobian row lists every column (src/nlpeval.jl:49-54)
    for i in 1:num_constr
        flatten!(op, arg, val, constr_body(MathProgBase.constr_expr(oracle, i))); push!(eptr, length(op))
        # KTN_ROW_NL = the rows optimize! tests every round: those loadproblem! keeps in nlconstr_ixs (src/model.jl:116-121, :148)
        flags[i] = (MathProgBase.isconstrlinear(oracle, i) ? 0x00 : 0x01) | (i == dense_row ? 0x02 : 0x00)
    end
    lb, ub = fill(-Inf, num_constr), fill(Inf, num_constr)     # real bounds arrive through set_bounds!
    check(sep, ccall((:ktn_add_rows, libktn), Cint,
          (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
          sep.handle, 0, num_constr, eptr, op, arg, val, lb, ub, flags), "ktn_add_rows")
    check(sep, ccall((:ktn_load_end, libktn), Cint, (Ptr{Cvoid},), sep.handle), "ktn_load_end")
    rp = Vector{Int64}(undef, num_constr + 1)
    check(sep, ccall((:ktn_jac_structure, libktn), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}), sep.handle, rp, C_NULL), "ktn_jac_structure")
    jc = Vector{Int32}(undef, rp[end])
    check(sep, ccall((:ktn_jac_structure, libktn), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int32}), sep.handle, rp, jc), "ktn_jac_structure")
    sep.cols = [Int.(jc[rp[i] + 1:rp[i + 1]]) .+ 1 for i in 1:num_constr]
    sep.linear_model = linear_model
    sep.num_var, sep.num_constr = num_var, num_constr
    sep.g = zeros(num_constr); sep.have_g = false
end

set_bounds!(sep::KatanaGPUSeparator, l, u) =
    check(sep, ccall((:ktn_set_bounds, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), sep.handle, l, u), "ktn_set_bounds")

# The batched round optimize! uses: returns (status, view).  status == 1 maps to m.status = :Error.
function separate!(sep::KatanaGPUSeparator, xstar::Vector{Float64})
    nc, nz, er = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    st = ccall((:ktn_separate, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), sep.handle, xstar, nc, nz, er)
    check(sep, st, "ktn_separate")
    v = Ref{KtnCutView}()
    check(sep, ccall((:ktn_fetch_cuts_view, libktn), Cint, (Ptr{Cvoid}, Ref{KtnCutView}), sep.handle, v), "ktn_fetch_cuts_view")
    sep.xstar = xstar; sep.have_g = false
    return st, v[]
end

# ---- the four reference hooks keep working for unmodified callers (per-row semantics, answered from the last round) ----
precompute!(sep::KatanaGPUSeparator, xstar) = (separate!(sep, xstar); nothing)                    # src/separators.jl:111
function isconstrsat(sep::KatanaGPUSeparator, i, lb, ub, f_tol)                                   # src/separators.jl:120
    if !sep.have_g
        check(sep, ccall((:ktn_get_g, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}), sep.handle, sep.g), "ktn_get_g"); sep.have_g = true
    end
    (sep.g[i] >= lb - f_tol) && (sep.g[i] <= ub + f_tol)
end
function gencut(sep::KatanaGPUSeparator, xstar, bounds, i)                                        # src/separators.jl:118
    rows = Int64[i - 1]; nc, nz, er = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    st = ccall((:ktn_gencut_rows, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Cint, Ref{Int64}, Ref{Int64}, Ref{Int64}),
               sep.handle, xstar, rows, 1, 0, nc, nz, er)
    check(sep, st, "ktn_gencut_rows")
    vars = [Variable(sep.linear_model, j) for j in sep.cols[i]]
    # a non-finite coefficient (status 1: the row yields no cut): hand _addcut NaN coefficients, so that ITS finiteness test sets
    # m.status = :Error and warns, exactly as with the reference separator (src/model.jl:69-73)
    (st == 1 || nc[] != 1) && return AffExpr(vars, fill(NaN, length(vars)), NaN)
    vals = Vector{Float64}(undef, nz[]); b = Vector{Float64}(undef, 1)     # a full (not lean) fetch: bconst is needed here
    check(sep, ccall((:ktn_fetch_cuts, libktn), Cint,
          (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
          sep.handle, C_NULL, C_NULL, C_NULL, vals, C_NULL, C_NULL, C_NULL, C_NULL, b), "ktn_fetch_cuts")
    AffExpr(vars, vals, b[1])
end

# boundroutine(m, ray) (src/model.jl:175-197) as ONE call: the points 2^n * ray, n = 2 .. 1023, are evaluated on the device in batches and the
# cuts are made at the first point that violates a row.  Plain handles only (a pipelined or multi-device handle keeps the reference's loop
# over separate!).  Returns (status, n_hit or -1, view).
function separate_ladder!(sep::KatanaGPUSeparator, ray::Vector{Float64}, n_first::Integer = 2, n_last::Integer = 1023)
    hit = Ref{Int32}(-1); nc, nz, er = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    st = ccall((:ktn_separate_ladder, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ref{Int32}, Ref{Int64}, Ref{Int64}, Ref{Int64}),
               sep.handle, ray, n_first, n_last, hit, nc, nz, er)
    check(sep, st, "ktn_separate_ladder")
    v = Ref{KtnCutView}()
    check(sep, ccall((:ktn_fetch_cuts_view, libktn), Cint, (Ptr{Cvoid}, Ref{KtnCutView}), sep.handle, v), "ktn_fetch_cuts_view")
    sep.xstar = (2.0^(hit[] >= 0 ? hit[] : n_last)) .* ray; sep.have_g = false
    return st, Int(hit[]), v[]
end

# loadproblem! (src/model.jl:172), right after `initialize!(sep, m.linear_model, m.num_var, m.num_constr, d)`:
set_bounds!(m.params.separator, m.l_constr, m.u_constr)   # bounds are model state (src/model.jl:273-277): handed over once

# optimize! (src/model.jl:265-283), instead of the per-row isconstrsat / gencut / round_coefs / _addcut loop:
xstar = MathProgBase.getsolution(mpb_lp)
st, v = separate!(m.params.separator, xstar)             # precompute! + test + cuts, ascending row order
colv = unsafe_wrap(Array, v.col, v.nnz); valv = unsafe_wrap(Array, v.val, v.nnz)
ptr  = unsafe_wrap(Array, v.row_ptr, v.n_cuts + 1)
lo   = unsafe_wrap(Array, v.lo, v.n_cuts); hi = unsafe_wrap(Array, v.hi, v.n_cuts)
for c in 1:v.n_cuts                                       # rows go to the LP straight from the pinned view
    r = (ptr[c] + 1):ptr[c + 1]
    MathProgBase.addconstr!(mpb_lp, colv[r] .+ 1, valv[r], lo[c], hi[c])   # [lb - b, ub - b], src/model.jl:74-75
end
m.numcuts += v.n_cuts
if st == 1                                                # non-finite coefficient in a selected row (src/model.jl:69-73)
    Base.warn("Nonlinear constraint or objective likely undefined within domain"); return m.status = :Error
end
allsat = v.n_cuts == 0

# Cut management (an extension: the reference never removes a cut, src/model.jl:215; mirrors katana.jl_b200/lp.py purge_slack_rows and the
# rule of model.py: purge only after the LP bound has passed its value at the last purge, put everything back if the LP loses its bound).
# `cut_rows` are the LP row indices of the loop cuts, `cut_lo` / `cut_hi` their bounds, `slack_age` how many optima in a row each has been
# slack.  Deletes the cuts slack for `age` consecutive optima and returns how many; the bookkeeping vectors are updated in place.
function purge_slack_cuts!(mpb_lp, cut_rows::Vector{Int}, cut_lo::Vector{Float64}, cut_hi::Vector{Float64}, slack_age::Vector{Int}, age::Int; tol = 1e-7)
    act = MathProgBase.getconstrsolution(mpb_lp)
    drop = Int[]
    for (k, r) in enumerate(cut_rows)
        slack = min(isfinite(cut_hi[k]) ? cut_hi[k] - act[r] : Inf, isfinite(cut_lo[k]) ? act[r] - cut_lo[k] : Inf)
        ref = max(1.0, isfinite(cut_hi[k]) ? abs(cut_hi[k]) : 0.0, isfinite(cut_lo[k]) ? abs(cut_lo[k]) : 0.0)
        slack_age[k] = slack > tol * ref ? slack_age[k] + 1 : 0
        slack_age[k] >= age && push!(drop, k)
    end
    isempty(drop) && return 0
    gone = sort(cut_rows[drop])
    MathProgBase.delconstrs!(mpb_lp, gone)
    deleteat!(cut_rows, drop); deleteat!(cut_lo, drop); deleteat!(cut_hi, drop); deleteat!(slack_age, drop)
    for k in eachindex(cut_rows); cut_rows[k] -= searchsortedlast(gone, cut_rows[k]); end      # the rows behind a deleted row move up
    return length(gone)
end

# Sharded rounds, one process per GPU (e.g. Distributed.jl workers): rank r of nranks owns rows [row_begin, row_end).
# (Inside ONE process, `KatanaGPUSeparator(ngpus = 8)` above needs none of this: ktn_options.ngpus.)
# `id` is the 128-byte communicator id made on rank 0 (ktn_comm_unique_id) and sent to the others by the host's own means.
function comm_unique_id()
    id = zeros(UInt8, 128)
    rc = ccall((:ktn_comm_unique_id, libktn), Cint, (Ptr{UInt8},), id)
    rc < 0 && error("ktn_comm_unique_id failed ($rc)")
    return id
end

function join_shards!(sep::KatanaGPUSeparator, nranks::Integer, rank::Integer, id::Vector{UInt8}, row_begin::Integer)
    check(sep, ccall((:ktn_set_row_offset, libktn), Cint, (Ptr{Cvoid}, Int64), sep.handle, row_begin), "ktn_set_row_offset")
    check(sep, ccall((:ktn_comm_init, libktn), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), sep.handle, nranks, rank, id), "ktn_comm_init")
end

# One sharded round: separate this rank's rows at x*, exchange, and return ALL ranks' cuts (global row ids, ascending) as CSR.
# The status is the same on every rank: 1 = some rank met a non-finite cut; the batch ends at that row (err_row, 1-based).
function separate_sharded(sep::KatanaGPUSeparator, xstar::Vector{Float64})
    nc = Ref{Int64}(0); nz = Ref{Int64}(0); er = Ref{Int64}(-1)
    check(sep, ccall((:ktn_separate, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), sep.handle, xstar, nc, nz, er), "ktn_separate")
    check(sep, ccall((:ktn_allgather_cuts_async, libktn), Cint, (Ptr{Cvoid},), sep.handle), "ktn_allgather_cuts_async")
    st = ccall((:ktn_sync_gathered, libktn), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), sep.handle, nc, nz)
    check(sep, st, "ktn_sync_gathered")
    check(sep, ccall((:ktn_gathered_error_row, libktn), Cint, (Ptr{Cvoid}, Ref{Int64}), sep.handle, er), "ktn_gathered_error_row")
    n, z = nc[], nz[]
    row = Vector{Int64}(undef, n); ptr = Vector{Int64}(undef, n + 1); col = Vector{Int32}(undef, z); val = Vector{Float64}(undef, z)
    lo = Vector{Float64}(undef, n); hi = Vector{Float64}(undef, n)
    check(sep, ccall((:ktn_fetch_gathered, libktn), Cint,
                     (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     sep.handle, row, ptr, col, val, lo, hi, C_NULL, C_NULL, C_NULL), "ktn_fetch_gathered")
    return st, er[] + 1, row .+ 1, ptr .+ 1, col .+ Int32(1), val, lo, hi          # 1-based for Julia
end
