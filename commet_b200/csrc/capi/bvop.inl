// stage 4: boolean-vector operators, popcounts; the measurement entry point
// (part of the C-ABI library: included by capi.cu, in this order, into one translation unit)

// ----------------------------------------------------------- stage 4: bvop --
extern "C" int commet_bvop_dev(commet_ctx *c, int op, const void *d_a, const void *d_b, void *d_out, uint64_t n_bytes)
{
    CKR(set_device(c));
    if (op < 0 || op > 3) return fail("unknown bv op %d", op);
    if (n_bytes == 0) return 0;
    if (((uintptr_t)d_a | (uintptr_t)d_out | (op == 3 ? 0 : (uintptr_t)d_b)) & 15) return fail("bvop buffers must be 16-byte aligned");
    uint64_t n_vec = n_bytes / 16;
    unsigned g = grid_for(c, std::max<uint64_t>(n_vec, 16), 256, 8);
    const uint4 *a = static_cast<const uint4 *>(d_a), *b = static_cast<const uint4 *>(d_b);
    uint4 *o = static_cast<uint4 *>(d_out);
    switch (op) {
    case 0: k_bvop<0><<<g, 256, 0, c->stream>>>(a, b, o, n_vec, n_bytes); break;
    case 1: k_bvop<1><<<g, 256, 0, c->stream>>>(a, b, o, n_vec, n_bytes); break;
    case 2: k_bvop<2><<<g, 256, 0, c->stream>>>(a, b, o, n_vec, n_bytes); break;
    default: k_bvop<3><<<g, 256, 0, c->stream>>>(a, a, o, n_vec, n_bytes); break;
    }
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// nb_one of several device-resident vectors: one kernel per vector, ONE read-back and ONE synchronisation for all
// (a count per call costs a D2H copy and a stream sync that dwarf the kernel: 125 MB are counted in 25 us)
extern "C" int commet_bv_popcount_batch_dev(commet_ctx *c, const void *const *d_bvs, const uint64_t *n_bits, int n, uint64_t *ones)
{
    CKR(set_device(c));
    if (n <= 0) return 0;
    if (!d_bvs || !n_bits || !ones) return fail("commet_bv_popcount_batch_dev: null argument");
    DevBuf tot(c);
    if (tot.alloc((size_t)n * sizeof(unsigned long long)) != cudaSuccess) return fail("popcount allocation failed");
    CK(cudaMemsetAsync(tot.p, 0, (size_t)n * sizeof(unsigned long long), c->stream));
    for (int i = 0; i < n; i++) {
        if ((uintptr_t)d_bvs[i] & 15) return fail("bv buffer must be 16-byte aligned");
        const uint64_t n_bytes = n_bits[i] / 8 + 1, n_vec = n_bytes / 16;
        k_popcount<<<grid_for(c, std::max<uint64_t>(n_vec, 16), 256, 8), 256, 0, c->stream>>>(static_cast<const uint4 *>(d_bvs[i]), n_vec,
                                                                                             n_bytes, tot.as<unsigned long long>() + i);
        c->launches++;
    }
    CK(cudaGetLastError());
    std::vector<unsigned long long> h(n);
    CK(cudaMemcpyAsync(h.data(), tot.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) ones[i] = h[i] > n_bits[i] ? n_bits[i] : h[i];      // boolean_vector.h:266-268
    return 0;
}

extern "C" int commet_bv_popcount_dev(commet_ctx *c, const void *d_bv, uint64_t n_bits, uint64_t *ones)
{
    uint64_t one = 0;
    CKR(commet_bv_popcount_batch_dev(c, &d_bv, &n_bits, 1, &one));
    if (ones) *ones = one;
    return 0;
}

extern "C" int commet_bvop(commet_ctx *c, int op, const uint8_t *a, const uint8_t *b, uint8_t *out, uint64_t n_bytes)
{
    CKR(set_device(c));
    if (n_bytes == 0) return 0;
    DevBuf da(c), db(c), dout(c);
    if (da.alloc(n_bytes) != cudaSuccess || dout.alloc(n_bytes) != cudaSuccess || (op != 3 && db.alloc(n_bytes) != cudaSuccess))
        return fail("bvop allocation failed");
    CK(cudaMemcpyAsync(da.p, a, n_bytes, cudaMemcpyHostToDevice, c->stream));
    if (op != 3) CK(cudaMemcpyAsync(db.p, b, n_bytes, cudaMemcpyHostToDevice, c->stream));
    CKR(commet_bvop_dev(c, op, da.p, db.p, dout.p, n_bytes));
    CK(cudaMemcpyAsync(out, dout.p, n_bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int commet_bv_popcount(commet_ctx *c, const uint8_t *bv, uint64_t n_bits, uint64_t *ones)
{
    CKR(set_device(c));
    uint64_t n_bytes = n_bits / 8 + 1;
    DevBuf d(c);
    if (d.alloc(n_bytes) != cudaSuccess) return fail("popcount allocation failed");
    CK(cudaMemcpyAsync(d.p, bv, n_bytes, cudaMemcpyHostToDevice, c->stream));
    return commet_bv_popcount_dev(c, d.p, n_bits, ones);
}

// ------------------------------------------------------------ measurement ---
extern "C" int commet_bench_random_sectors(commet_ctx *c, uint64_t bytes, uint64_t n_ops, int atomic_or, double *ns)
{
    CKR(set_device(c));
    if (bytes < 4096 || (bytes & (bytes - 1))) return fail("bytes must be a power of two >= 4096");
    DevBuf buf(c);
    if (buf.alloc(bytes) != cudaSuccess) return fail("allocation of %llu bytes failed", (unsigned long long)bytes);
    CK(cudaMemsetAsync(buf.p, 0, bytes, c->stream));
    uint64_t mask = bytes / 4 - 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    unsigned g = grid_for(c, n_ops / 4, 256, 64);  // several waves: see grid_for
    for (int rep = 0; rep < 2; rep++) {          // first pass warms up, second is timed
        if (rep == 1) CK(cudaEventRecord(e0, c->stream));
        if (atomic_or) k_random_sectors<true><<<g, 256, 0, c->stream>>>(buf.as<uint32_t>(), mask, n_ops, c->scratch + 144);
        else k_random_sectors<false><<<g, 256, 0, c->stream>>>(buf.as<uint32_t>(), mask, n_ops, c->scratch + 144);
        c->launches++;
    }
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ns) *ns = (double)ms * 1e6;
    return 0;
}
