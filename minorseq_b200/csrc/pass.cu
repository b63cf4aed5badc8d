// pass.cu -- the whole juliet pass behind one C-ABI call: pileup -> (all-reduce) -> codon test -> phasing.
//
// What `juliet [--mode-phasing] in.bam out.json` does between BAM decode and report writing
// (/root/reference/doc/JULIET.md:38-42, :192-211), as one entry point so that a host program pays two small
// device->host reads per pass instead of a round trip per stage.  Composition only: every stage is the same
// code as the stand-alone entry points.
#include <algorithm>
#include <cstring>
#include <vector>
#include "handle.h"
#include "phase_internal.cuh"

enum RowSource { kDeviceRows, kHostRows, kHostEvents };

int ms_call_launch(ms_handle* h, const ms_gene* genes, int32_t ngenes, const char* refseq, const ms_call_params* prm,
                   const ms_variant** d_calls, const unsigned long long** d_ncalls, int64_t* calls_cap);
int ms_call_collect(ms_handle* h, bool synced, ms_variant* out, int64_t cap, int64_t* n);
int ms_phase_planned_dev(ms_handle* h, const ms_variant* d_calls, const unsigned long long* d_ncalls, int64_t calls_cap,
                         const uint32_t* d_packed, int64_t R, ms::PhasePlan** d_plan_out);
void ms_phase_planned_adopt(ms_handle* h, int32_t V);

static int juliet_pass_body(ms_handle* h, const void* src, const uint8_t* events, RowSource from, int64_t R, const ms_gene* genes, int32_t ngenes,
                            const char* refseq, const ms_call_params* prm, int32_t phase, int32_t min_hap_reads, ms_juliet_result* out);

static int juliet_pass(ms_handle* h, const void* src, const uint8_t* events, RowSource from, int64_t R, const ms_gene* genes, int32_t ngenes,
                       const char* refseq, const ms_call_params* prm, int32_t phase, int32_t min_hap_reads, ms_juliet_result* out) {
    if (!h || !out || !prm || R < 0) return MS_ERR_ARG;
    MS_CUDA(h, cudaSetDevice(h->device));
    MS_STAGE_BEGIN(h, MS_STAGE_PASS);
    int rc = juliet_pass_body(h, src, events, from, R, genes, ngenes, refseq, prm, phase, min_hap_reads, out);
    MS_STAGE_END(h, MS_STAGE_PASS);       // (the body has waited for its last result: this is when it arrived)
    return rc;
}

static int juliet_pass_body(ms_handle* h, const void* src, const uint8_t* events, RowSource from, int64_t R, const ms_gene* genes, int32_t ngenes,
                            const char* refseq, const ms_call_params* prm, int32_t phase, int32_t min_hap_reads, ms_juliet_result* out) {
    int rc = ms_reset_counts(h);
    if (rc != MS_OK) return rc;
    const uint32_t* d_rows = static_cast<const uint32_t*>(src);
    if (from == kHostRows) rc = ms_pileup_host(h, static_cast<const uint32_t*>(src), R, &d_rows);
    else if (from == kHostEvents) rc = ms_pileup_events_host(h, static_cast<const ms_read_hdr*>(src), events, R, &d_rows);
    else rc = ms_pileup_dev(h, d_rows, R);
    if (rc != MS_OK) return rc;
    rc = ms_allreduce_counts(h);
    if (rc != MS_OK) return rc;
    out->nkeys = 0; out->npatterns = 0; out->nreported = 0; out->nvariants = 0;
    memset(&out->counters, 0, sizeof out->counters);
    const ms_variant* d_calls = nullptr;
    const unsigned long long* d_ncalls = nullptr;
    int64_t calls_cap = 0;
    rc = ms_call_launch(h, genes, ngenes, refseq, prm, &d_calls, &d_ncalls, &calls_cap);
    if (rc != MS_OK) return rc;
    if (!phase) return ms_call_collect(h, false, out->variants, out->variants_cap, &out->nvariants);

    // Phasing without a host round trip in the middle of the pass: the plan (pooled variant list, touched blocks, the
    // bit-vector kernel's word stream) is built on the device from K2's output, the bit-vectors, the grouping and the order
    // follow on the stream, and calls, plan and haplotypes come down together at the end.  The plan covers the usual case
    // (<= 32 distinct variant codons); when it gives up, or a caller buffer is too small, the host-planned path below runs.
    static const bool planner_on = getenv("MS_NO_PLANNER") == nullptr;
    if (planner_on && d_calls && out->patterns_cap > 0 && out->keys_cap >= ms::kPlanMaxKeys) {
        ms::PhasePlan* d_plan = nullptr;
        rc = ms_phase_planned_dev(h, d_calls, d_ncalls, calls_cap, d_rows, R, &d_plan);
        if (rc != MS_OK) return rc;
        if (!h->plan_stage) MS_CUDA(h, cudaMallocHost(&h->plan_stage, sizeof(ms::PhasePlan)));
        MS_CUDA(h, cudaMemcpyAsync(h->plan_stage, d_plan, sizeof(ms::PhasePlan), cudaMemcpyDeviceToHost, h->stream));
        int64_t H = 0, nrep = 0;
        rc = ms_phase_haplotypes(h, min_hap_reads, out->patterns, out->counts, out->patterns_cap, &H, &nrep, &out->counters, out->hap_id);
        if (rc != MS_OK) return rc;          // (synchronises the stream: the calls and the plan have arrived as well)
        ms::PhasePlan plan;
        memcpy(&plan, h->plan_stage, sizeof plan);
        rc = ms_call_collect(h, true, out->variants, out->variants_cap, &out->nvariants);
        if (rc != MS_OK) return rc;
        if (!plan.fallback) {
            if (out->nvariants > out->variants_cap) return MS_ERR_CAPACITY;   // the caller re-runs with a larger buffer
            ms_phase_planned_adopt(h, plan.V);
            out->nkeys = plan.V;
            for (int32_t i = 0; i < plan.V; ++i) { out->key_col[i] = plan.key_col[i]; out->key_codon[i] = plan.key_codon[i]; }
            out->npatterns = H;
            out->nreported = nrep;
            if (nrep > out->patterns_cap) return MS_ERR_CAPACITY;
            return MS_OK;
        }
        memset(&out->counters, 0, sizeof out->counters);
    } else {
        rc = ms_call_collect(h, false, out->variants, out->variants_cap, &out->nvariants);
        if (rc != MS_OK) return rc;
    }
    if (out->nvariants > out->variants_cap) return MS_ERR_CAPACITY;   // the caller re-runs with a larger buffer
    // one pooled, de-duplicated variant list over all genes (screenshot juliet_hiv-phasing.png)
    std::vector<std::pair<int32_t, int32_t>> keys;
    keys.reserve(static_cast<size_t>(out->nvariants));
    for (int64_t i = 0; i < out->nvariants; ++i) keys.emplace_back(out->variants[i].col, out->variants[i].codon);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const int32_t V = static_cast<int32_t>(keys.size());
    out->nkeys = V;
    if (V > out->keys_cap) return MS_ERR_CAPACITY;
    std::vector<int32_t> vc(std::max(1, V)), vk(std::max(1, V));
    for (int32_t i = 0; i < V; ++i) { vc[i] = keys[i].first; vk[i] = keys[i].second; out->key_col[i] = vc[i]; out->key_codon[i] = vk[i]; }
    rc = ms_phase_begin(h, vc.data(), vk.data(), V, R);
    if (rc != MS_OK) return rc;
    rc = ms_phase_dev(h, d_rows, R);
    if (rc != MS_OK) return rc;
    int64_t H = 0, nrep = 0;
    rc = ms_phase_haplotypes(h, min_hap_reads, out->patterns, out->counts, out->patterns_cap, &H, &nrep, &out->counters, out->hap_id);
    if (rc != MS_OK) return rc;
    out->npatterns = H;
    out->nreported = nrep;
    if (nrep > out->patterns_cap) return MS_ERR_CAPACITY;
    return MS_OK;
}

extern "C" {

int ms_juliet_pass_dev(ms_handle* h, const uint32_t* d_packed, int64_t R, const ms_gene* genes, int32_t ngenes, const char* refseq,
                       const ms_call_params* prm, int32_t phase, int32_t min_hap_reads, ms_juliet_result* out) {
    return juliet_pass(h, d_packed, nullptr, kDeviceRows, R, genes, ngenes, refseq, prm, phase, min_hap_reads, out);
}

int ms_juliet_pass_host(ms_handle* h, const uint32_t* h_packed, int64_t R, const ms_gene* genes, int32_t ngenes, const char* refseq,
                        const ms_call_params* prm, int32_t phase, int32_t min_hap_reads, ms_juliet_result* out) {
    return juliet_pass(h, h_packed, nullptr, kHostRows, R, genes, ngenes, refseq, prm, phase, min_hap_reads, out);
}

int ms_juliet_pass_events_host(ms_handle* h, const ms_read_hdr* hdr, const uint8_t* events, int64_t R, const ms_gene* genes, int32_t ngenes,
                               const char* refseq, const ms_call_params* prm, int32_t phase, int32_t min_hap_reads, ms_juliet_result* out) {
    return juliet_pass(h, hdr, events, kHostEvents, R, genes, ngenes, refseq, prm, phase, min_hap_reads, out);
}

}  // extern "C"
