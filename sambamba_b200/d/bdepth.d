/**
  bdepth.d -- D `extern(C)` binding of libbdepth.so (include/bdepth.h) and a sketch of the patch to
  sambamba/depth.d that routes `sambamba depth` through it.

  NOT COMPILED IN THIS REPOSITORY'S CI: the authoring image has no D compiler (no ldc2/dmd/gdc; SURVEY.md F1).
  The declarations are a line-for-line transcription of include/bdepth.h; the C++ host in
  sambamba_b200/csrc/cli.cpp is the tested equivalent of `gpuDepthMain` below.

  Build (where LDC exists):  ldc2 -O3 -release bdepth.d depth_gpu_main.d -L-L<repo>/sambamba_b200/_build -L-lbdepth
*/
module bdepth;

extern (C) nothrow @nogc:

struct bdepth_t;   // opaque

enum : int {
    BDEPTH_OK = 0, BDEPTH_ERR_IO = -1, BDEPTH_ERR_FORMAT = -2, BDEPTH_ERR_UNSORTED = -3, BDEPTH_ERR_NOINDEX = -4,
    BDEPTH_ERR_CUDA = -5, BDEPTH_ERR_NCCL = -6, BDEPTH_ERR_ARG = -7, BDEPTH_ERR_CALLBACK = -8
}

/// same meaning as bio.std.hts.bam.region.BamRegion (0-based, half open)
struct bdepth_region { uint ref_id, start, end; }

/// counts is SoA: plane p (A,C,G,T,N,DEL,REFSKIP) at counts[p*stride .. p*stride+len]
struct bdepth_tile { int ref_id; uint start, len, stride; const(uint)* counts; uint n_samples, sample_stride; }
alias bdepth_tile_cb = int function(void* user, const(bdepth_tile)* tile);

struct bdepth_region_stat { int ref_id; uint start, end, n_reads, n_bases; const(uint)* cov_ge; int sample_id; }
alias bdepth_stat_cb = int function(void* user, const(bdepth_region_stat)* s, ulong index);

struct bdepth_stats {
    ulong file_bytes, n_blocks, cdata_bytes, inflated_bytes, n_records, n_records_pass, n_cigar_ops, seq_bytes,
          positions, covered_positions, long_reads, chain_fixups;
    uint gpu_launches, n_batches;
    float ms_h2d, ms_inflate, ms_scan, ms_coverage, ms_reduce, ms_d2h, ms_total_device;
    double host_wall_ms;
    float ms_span_device, ms_exchange;
    ulong own_lo, own_hi, halo_bytes_sent;
    ulong mate_pairs, mate_pair_columns, mate_groups;
    float ms_mates;
}

int bdepth_device_count();
int bdepth_open(const(char)* bam_path, int device, bdepth_t** h);
int bdepth_open_lazy(const(char)* bam_path, int device, bdepth_t** h);   // region queries: frame only what the BAI chunks touch
int bdepth_open_memory(const(void)* bam, size_t bam_len, const(void)* bai, size_t bai_len, int device, bdepth_t** h);
int bdepth_add_input(bdepth_t* h, const(char)* bam_path);
void bdepth_close(bdepth_t* h);
const(char)* bdepth_last_error(const(bdepth_t)* h);

int bdepth_n_ref(const(bdepth_t)* h);
const(char)* bdepth_ref_name(const(bdepth_t)* h, int i);
uint bdepth_ref_length(const(bdepth_t)* h, int i);
const(char)* bdepth_header_text(const(bdepth_t)* h, size_t* len);
int bdepth_is_coordinate_sorted(const(bdepth_t)* h);
int bdepth_has_index(const(bdepth_t)* h);
int bdepth_n_samples(const(bdepth_t)* h);
const(char)* bdepth_sample_name(const(bdepth_t)* h, int i);

int bdepth_set_filter(bdepth_t* h, int mapq_gt, uint flag_reject_mask);
int bdepth_set_filter_query(bdepth_t* h, const(char)* query);   // -F, depth.d:1121
int bdepth_set_min_baseq(bdepth_t* h, uint q);
int bdepth_set_fix_mates(bdepth_t* h, int on);   // -m, depth.d:1133
int bdepth_set_combined(bdepth_t* h, int combined);
int bdepth_set_regions(bdepth_t* h, const(bdepth_region)* r, size_t n);
int bdepth_set_shard(bdepth_t* h, int rank, int world, const(void)* nccl_unique_id);
int bdepth_nccl_unique_id(void* out128);
int bdepth_plan_shards(const(char)* bam_path, int world, ulong* out_voffsets);
long bdepth_plan_region_chunks(const(char)* bam_path, const(bdepth_region)* regions, size_t n, ulong* out_pairs, size_t cap);
int bdepth_set_tuning(bdepth_t* h, ulong batch_inflated_bytes, ulong chunk_blocks);

int bdepth_stage(bdepth_t* h);
int bdepth_run_resident(bdepth_t* h);
int bdepth_run_base(bdepth_t* h, bdepth_tile_cb cb, void* user);
int bdepth_run_windows(bdepth_t* h, uint window, uint overlap, const(uint)* thr, size_t n_thr, bdepth_stat_cb cb, void* user);
int bdepth_run_regions(bdepth_t* h, const(bdepth_region)* r, size_t n, const(uint)* thr, size_t n_thr, bdepth_stat_cb cb, void* user);
struct bdepth_text_opts { double min_cov, max_cov; int annotate; }
alias bdepth_text_cb = int function(void* user, const(char)* text, size_t len);
int bdepth_run_base_text(bdepth_t* h, const(bdepth_text_opts)* o, bdepth_text_cb cb, void* user);
int bdepth_get_stats(const(bdepth_t)* h, bdepth_stats* s);
int bdepth_ref_has_reads(const(bdepth_t)* h, int r);
long bdepth_inflate_to_host(bdepth_t* h, void* dst, ulong cap);
long bdepth_scan_to_host(bdepth_t* h, ulong cap, int* ref_id, int* pos, uint* span, ushort* flag, ubyte* mapq, ushort* n_cigar, ulong* rec_off);
/// createIndex (bio/std/hts/bam/bai/indexing.d:356) on the GPU; the handle adopts the index
long bdepth_build_index(bdepth_t* h, void* dst, ulong cap);

/+ ---------------------------------------------------------------------------------------------------------
   Sketch of the sambamba-side patch (sambamba/depth.d).  The option parsing of depth_main (depth.d:1121-1152)
   and the printers' text formats stay as they are; only the column source changes:

     // depth.d:1163-1234, replaced when a GPU is present
     bdepth_t* h;
     enforce(bdepth_open(bam_filenames[0].toStringz, 0, &h) == 0, bdepth_last_error(null).fromStringz);
     scope(exit) bdepth_close(h);
     enforce(bdepth_is_coordinate_sorted(h), "All files must be coordinate-sorted");     // depth.d:1164
     enforce(bdepth_has_index(h), "All files must be indexed");                           // depth.d:1166
     bdepth_set_filter(h, 0, 0x600);              // default filter, depth.d:1159
     bdepth_set_min_baseq(h, printer.min_base_quality);
     bdepth_set_fix_mates(h, printer.fix_mate_overlaps);                                   // depth.d:1133
     final switch (mode) {
       case Mode.base:   bdepth_run_base(h, &onTile, cast(void*) printer);   break;  // PerBasePrinter rows from tile planes
       case Mode.window: bdepth_run_windows(h, w, overlap, thr.ptr, thr.length, &onStat, cast(void*) printer); break;
       case Mode.region: bdepth_run_regions(h, bed.ptr, bed.length, thr.ptr, thr.length, &onStat, cast(void*) printer); break;
     }

   onTile walks `tile.len` positions and prints `REF POS COV A C G T DEL REFSKIP[ SAMPLE][ FLAG]` exactly as
   PerBasePrinter.writeColumn does (depth.d:534-555) with COV = sum of the seven planes; onStat feeds
   printRegionStats (depth.d:847-876) with n_reads / n_bases / cov_ge[].
+/
