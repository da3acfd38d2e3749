# gpu_separator.jl -- the reference-side binding of the B200 separation library (libktn.so, include/ktn.h).
#
# A Katana.jl maintainer adds this file to src/ and `include`s it from src/Katana.jl after separators.jl.  It is the code of
# INTEGRATION.md sections 3, 4 and 6 verbatim (tests/test_abi.py keeps the two in step and checks every ccall name against the
# exported symbols).  NOT executed in this repository: the build image has no julia binary.
#
mutable struct KatanaGPUSeparator <: AbstractKatanaSeparator
    handle :: Ptr{Cvoid}
    num_var :: Int
    num_constr :: Int
    xstar :: Vector{Float64}
    g :: Vector{Float64}            # filled lazily for the per-row isconstrsat hook
    have_g :: Bool
    KatanaGPUSeparator() = new(C_NULL, 0, 0, Float64[], Float64[], false)
end

struct KtnOptions            # include/ktn.h: ktn_options
    struct_size::Int32; device::Int32; f_tol::Float64; cut_coef_rng::Float64; topk::Int64; flags::Int32; reserved::Int32
end

struct KtnCutView            # include/ktn.h: ktn_cut_view
    n_cuts::Int64; nnz::Int64
    row_id::Ptr{Int64}; row_ptr::Ptr{Int64}; col::Ptr{Int32}; val::Ptr{Float64}
    lo::Ptr{Float64}; hi::Ptr{Float64}; g::Ptr{Float64}; viol::Ptr{Float64}; bconst::Ptr{Float64}
end

check(sep, rc, what) = rc < 0 && error("$what failed ($rc): " *
    unsafe_string(ccall((:ktn_last_error, libktn), Cstring, (Ptr{Cvoid},), sep.handle)))

# op codes of the wire format (include/ktn.h KTN_OP_*)
const KTN_OPS = Dict(:+ => 2, :- => 3, :* => 4, :/ => 5, :^ => 6, :exp => 8, :log => 9, :sqrt => 10, :abs => 11)

# Flatten one MathProgBase constraint expression `lhs <= / >= / == rhs` (an Expr tree over x[i]) into prefix arrays.
function flatten!(op::Vector{Int32}, arg::Vector{Int32}, val::Vector{Float64}, ex)
    if ex isa Number
        push!(op, 0); push!(arg, 0); push!(val, Float64(ex))
    elseif ex isa Expr && ex.head == :ref                     # x[i]
        push!(op, 1); push!(arg, Int32(ex.args[2] - 1)); push!(val, 0.0)
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

# initialize!(sep, linear_model, num_var, num_constr, oracle)           -- src/separators.jl:81-107
function initialize!(sep::KatanaGPUSeparator, linear_model, num_var::Int, num_constr::Int, oracle)
    MathProgBase.initialize(oracle, [:ExprGraph])              # the separator initialises the oracle itself (:88)
    if sep.handle == C_NULL                                    # one separator is reused across models (test/runtests.jl:24)
        o = Ref(KtnOptions(sizeof(KtnOptions), -1, 1e-6, 1e9, 0, 0, 0)); h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ktn_create, libktn), Cint, (Ref{KtnOptions}, Ref{Ptr{Cvoid}}), o, h)
        rc == 0 || error("ktn_create failed ($rc): no B200 / CUDA device?")
        sep.handle = h[]
    end
    check(sep, ccall((:ktn_load_begin, libktn), Cint, (Ptr{Cvoid}, Int64, Int64), sep.handle, num_var, num_constr), "ktn_load_begin")
    op, arg, val, eptr = Int32[], Int32[], Float64[], Int64[0]
    for i in 1:num_constr
        c = MathProgBase.constr_expr(oracle, i)                # :(lhs <= rhs) etc.; only lhs is compiled, bounds travel separately
        flatten!(op, arg, val, c.args[2]); push!(eptr, length(op))
    end
    lb, ub = fill(-Inf, num_constr), fill(Inf, num_constr)     # real bounds arrive through set_bounds!
    flags = fill(UInt8(1), num_constr)                         # KTN_ROW_NL; the epigraph row (nlpeval.jl) adds KTN_ROW_DENSE = 2
    check(sep, ccall((:ktn_add_rows, libktn), Cint,
          (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}),
          sep.handle, 0, num_constr, eptr, op, arg, val, lb, ub, flags), "ktn_add_rows")
    check(sep, ccall((:ktn_load_end, libktn), Cint, (Ptr{Cvoid},), sep.handle), "ktn_load_end")
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
    check(sep, ccall((:ktn_gencut_rows, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Cint, Ref{Int64}, Ref{Int64}, Ref{Int64}),
          sep.handle, xstar, rows, 1, 0, nc, nz, er), "ktn_gencut_rows")
    v = Ref{KtnCutView}(); ccall((:ktn_fetch_cuts_view, libktn), Cint, (Ptr{Cvoid}, Ref{KtnCutView}), sep.handle, v)
    cols = unsafe_wrap(Array, v[].col, v[].nnz) .+ 1; vals = unsafe_wrap(Array, v[].val, v[].nnz)
    AffExpr([Variable(sep.linear_model, j) for j in cols], copy(vals), unsafe_load(v[].bconst))
end

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

# Sharded rounds: one process per GPU (e.g. Distributed.jl workers), rank r of nranks owns rows [row_begin, row_end).
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
function separate_sharded(sep::KatanaGPUSeparator, xstar::Vector{Float64})
    nc = Ref{Int64}(0); nz = Ref{Int64}(0); er = Ref{Int64}(-1)
    st = ccall((:ktn_separate, libktn), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), sep.handle, xstar, nc, nz, er)
    check(sep, st, "ktn_separate")
    check(sep, ccall((:ktn_allgather_cuts_async, libktn), Cint, (Ptr{Cvoid},), sep.handle), "ktn_allgather_cuts_async")
    check(sep, ccall((:ktn_sync_gathered, libktn), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), sep.handle, nc, nz), "ktn_sync_gathered")
    n, z = nc[], nz[]
    row = Vector{Int64}(undef, n); ptr = Vector{Int64}(undef, n + 1); col = Vector{Int32}(undef, z); val = Vector{Float64}(undef, z)
    lo = Vector{Float64}(undef, n); hi = Vector{Float64}(undef, n)
    check(sep, ccall((:ktn_fetch_gathered, libktn), Cint,
                     (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                     sep.handle, row, ptr, col, val, lo, hi, C_NULL, C_NULL, C_NULL), "ktn_fetch_gathered")
    return st, row .+ 1, ptr .+ 1, col .+ Int32(1), val, lo, hi          # 1-based for Julia; st == 1: a rank met a non-finite cut (:Error)
end
