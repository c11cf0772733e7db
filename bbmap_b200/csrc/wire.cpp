// wire.cpp — the on-disk formats of a BBMap index directory, host side, no device code.
//
// The reference persists its index with java.io.ObjectOutputStream (current/fileIO/ReadWrite.java:200-240, readObject :739-757):
//   ref/index/<build>/chr<a>[-<b>]_index_k<k>_c<chrombits>_b<build>.block      int[] sites                 (Block.write, align2/Block.java:74-112; name: IndexMaker4.java:477-488)
//   ...block2.gz                                                               int[] starts, delta-coded (x[i]=starts[i]-starts[i-1], i>=1; Block.java:114-118), gzip
//   ref/genome/<build>/chr<N>.chrom.gz                                         dna.ChromosomeArray object (dna/ChromosomeArray.java:14-22,415-419), gzip
//   ref/genome/<build>/summary.txt                                             "key\tvalue" lines (dna/FastaToChromArrays2.java:229-250,371-403)
// A file whose name ends in .gz goes through gzip (ReadWrite.getOutputStream :349-392); everything else is written raw.
//
// Java Object Serialization Stream Protocol (version 5), as far as these files use it:
//   stream      := 0xACED 0x0005 content
//   int[]       := TC_ARRAY(0x75) classDesc("[I", suid 0x4DBA602676EAB2A5, SC_SERIALIZABLE, 0 fields) int32 length, length x int32 big-endian
//   byte[]      := TC_ARRAY classDesc("[B", suid 0xACF317F8060854E0, ...) int32 length, bytes
//   classDesc   := TC_CLASSDESC(0x72) utf(name) int64 suid, flags(0x02) int16 nfields field* TC_ENDBLOCKDATA(0x78) superclass(TC_NULL 0x70)
//   field       := typecode utf(name) [TC_STRING(0x74) utf(type signature) for '[' and 'L']
//   object      := TC_OBJECT(0x73) classDesc, then the field values in descriptor order (primitives first, sorted by name; then object fields by name)
// dna.ChromosomeArray (suid 3199182397853127842): int chromosome, int maxIndex, int minIndex, byte strand, byte[] array.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>
#include "../../include/bbmap_cuda.h"

namespace {

thread_local std::string g_wire_err;
int wfail(int code, const std::string& what) { g_wire_err = what; return code; }

bool ends_with(const char* s, const char* suf) { const size_t n = strlen(s), m = strlen(suf); return n >= m && memcmp(s + n - m, suf, m) == 0; }

// ---- byte sink / source over a raw FILE* or a gzFile, chosen by the file name like ReadWrite.getOutputStream / getInputStream ----
struct Sink {
    FILE* f = nullptr; gzFile g = nullptr; bool ok = true;
    bool open(const char* path) {
        if (ends_with(path, ".gz") || ends_with(path, ".gzip")) { g = gzopen(path, "wb4"); if (g) gzbuffer(g, 1 << 20); return g != nullptr; }
        f = fopen(path, "wb"); return f != nullptr;
    }
    void put(const void* p, size_t n) {
        if (!ok || n == 0) return;
        if (g) { const char* q = (const char*)p; while (n) { const unsigned c = (unsigned)std::min<size_t>(n, 1u << 30); if (gzwrite(g, q, c) != (int)c) { ok = false; return; } q += c; n -= c; } }
        else if (fwrite(p, 1, n, f) != n) ok = false;
    }
    void u8(unsigned v) { const unsigned char b = (unsigned char)v; put(&b, 1); }
    void u16(unsigned v) { const unsigned char b[2] = {(unsigned char)(v >> 8), (unsigned char)v}; put(b, 2); }
    void u32(uint32_t v) { const unsigned char b[4] = {(unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v}; put(b, 4); }
    void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
    void utf(const char* s) { const size_t n = strlen(s); u16((unsigned)n); put(s, n); }
    bool close() { bool r = ok; if (g) { r = (gzclose(g) == Z_OK) && r; g = nullptr; } if (f) { r = (fclose(f) == 0) && r; f = nullptr; } return r; }
    ~Sink() { close(); }
};

struct Source {
    gzFile g = nullptr; bool ok = true;        // gzread is transparent for files that are not gzip streams (.block)
    bool open(const char* path) { g = gzopen(path, "rb"); if (g) gzbuffer(g, 1 << 20); return g != nullptr; }
    void get(void* p, size_t n) {
        char* q = (char*)p;
        while (ok && n) { const unsigned c = (unsigned)std::min<size_t>(n, 1u << 30); const int r = gzread(g, q, c); if (r != (int)c) { ok = false; return; } q += c; n -= c; }
    }
    unsigned u8() { unsigned char b = 0; get(&b, 1); return b; }
    unsigned u16() { unsigned char b[2] = {0, 0}; get(b, 2); return (b[0] << 8) | b[1]; }
    uint32_t u32() { unsigned char b[4] = {0, 0, 0, 0}; get(b, 4); return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }
    uint64_t u64() { const uint64_t h = u32(); return (h << 32) | u32(); }
    std::string utf() { const unsigned n = u16(); std::string s(n, '\0'); if (n) get(&s[0], n); return s; }
    ~Source() { if (g) gzclose(g); }
};

constexpr unsigned TC_NULL = 0x70, TC_CLASSDESC = 0x72, TC_OBJECT = 0x73, TC_STRING = 0x74, TC_ARRAY = 0x75, TC_ENDBLOCKDATA = 0x78;
constexpr unsigned SC_SERIALIZABLE = 0x02;
constexpr uint64_t SUID_INT_ARRAY = 0x4DBA602676EAB2A5ull, SUID_BYTE_ARRAY = 0xACF317F8060854E0ull, SUID_CHROMOSOME_ARRAY = 3199182397853127842ull;

void put_header(Sink& s) { s.u16(0xACED); s.u16(5); }
void put_array_desc(Sink& s, const char* name, uint64_t suid) {
    s.u8(TC_ARRAY); s.u8(TC_CLASSDESC); s.utf(name); s.u64(suid); s.u8(SC_SERIALIZABLE); s.u16(0); s.u8(TC_ENDBLOCKDATA); s.u8(TC_NULL);
}
void put_ints_be(Sink& s, const int32_t* d, int64_t n) {
    std::vector<uint32_t> buf(1 << 16);
    for (int64_t i = 0; i < n;) {
        const int64_t c = std::min<int64_t>(n - i, (int64_t)buf.size());
        for (int64_t j = 0; j < c; ++j) buf[j] = __builtin_bswap32((uint32_t)d[i + j]);
        s.put(buf.data(), (size_t)c * 4); i += c;
    }
}

struct FieldDesc { char type; std::string name, sig; };
struct ClassDesc { std::string name; uint64_t suid = 0; unsigned flags = 0; std::vector<FieldDesc> fields; };

// classDesc as it appears first in a stream (no back references: each of these files holds one object graph whose classes occur once)
int get_class_desc(Source& s, ClassDesc& d, const char* path) {
    const unsigned tc = s.u8();
    if (tc != TC_CLASSDESC) return wfail(BBM_E_ARG, std::string(path) + ": expected TC_CLASSDESC, found 0x" + std::to_string(tc));
    d.name = s.utf(); d.suid = s.u64(); d.flags = s.u8();
    const unsigned nf = s.u16();
    for (unsigned i = 0; i < nf && s.ok; ++i) {
        FieldDesc f; f.type = (char)s.u8(); f.name = s.utf();
        if (f.type == '[' || f.type == 'L') { if (s.u8() != TC_STRING) return wfail(BBM_E_ARG, std::string(path) + ": field type signature is not a new string"); f.sig = s.utf(); }
        d.fields.push_back(f);
    }
    if (s.u8() != TC_ENDBLOCKDATA) return wfail(BBM_E_ARG, std::string(path) + ": class annotation present (not written by default serialization)");
    if (s.u8() != TC_NULL) return wfail(BBM_E_ARG, std::string(path) + ": class has a serializable superclass");
    return s.ok ? 0 : wfail(BBM_E_ARG, std::string(path) + ": truncated class descriptor");
}
int get_header(Source& s, const char* path) {
    if (s.u16() != 0xACED || s.u16() != 5 || !s.ok) return wfail(BBM_E_ARG, std::string(path) + ": not a Java object stream (magic ACED0005)");
    return 0;
}
int get_array_head(Source& s, const char* name, uint64_t suid, int64_t* n, const char* path) {
    if (s.u8() != TC_ARRAY) return wfail(BBM_E_ARG, std::string(path) + ": expected an array");
    ClassDesc d; if (int e = get_class_desc(s, d, path)) return e;
    if (d.name != name || d.suid != suid || !d.fields.empty()) return wfail(BBM_E_ARG, std::string(path) + ": array class is " + d.name + ", expected " + name);
    const int32_t len = (int32_t)s.u32();
    if (!s.ok || len < 0) return wfail(BBM_E_ARG, std::string(path) + ": bad array length");
    *n = len; return 0;
}

}  // namespace

extern "C" const char* bbm_wire_last_error() { return g_wire_err.c_str(); }
extern "C" void bbm_wire_free(void* p) { free(p); }

// ReadWrite.write(int[] x, fname): ObjectOutputStream.writeObject of an int array
extern "C" int bbm_wire_write_int_array(const char* path, const int32_t* data, int64_t n) {
    if (!path || n < 0 || n > 0x7fffffffLL || (n && !data)) return wfail(BBM_E_ARG, "bbm_wire_write_int_array: bad argument");
    Sink s; if (!s.open(path)) return wfail(BBM_E_ARG, std::string("cannot create ") + path);
    put_header(s); put_array_desc(s, "[I", SUID_INT_ARRAY); s.u32((uint32_t)n); put_ints_be(s, data, n);
    return s.close() ? BBM_OK : wfail(BBM_E_ARG, std::string("write failed: ") + path);
}

// ReadWrite.read(int[].class, fname)
extern "C" int bbm_wire_read_int_array(const char* path, int32_t** data_out, int64_t* n_out) {
    if (!path || !data_out || !n_out) return wfail(BBM_E_ARG, "bbm_wire_read_int_array: bad argument");
    Source s; if (!s.open(path)) return wfail(BBM_E_ARG, std::string("cannot open ") + path);
    if (int e = get_header(s, path)) return e;
    int64_t n = 0; if (int e = get_array_head(s, "[I", SUID_INT_ARRAY, &n, path)) return e;
    int32_t* d = (int32_t*)malloc((size_t)std::max<int64_t>(n, 1) * 4);
    if (!d) return wfail(BBM_E_ARG, "out of memory");
    s.get(d, (size_t)n * 4);
    if (!s.ok) { free(d); return wfail(BBM_E_ARG, std::string(path) + ": truncated array data"); }
    for (int64_t i = 0; i < n; ++i) d[i] = (int32_t)__builtin_bswap32((uint32_t)d[i]);
    *data_out = d; *n_out = n; return BBM_OK;
}

// IndexMaker4.fname (IndexMaker4.java:477-488)
extern "C" int bbm_wire_block_fname(char* out, size_t cap, const char* root_index, int minChrom, int maxChrom, int k, int chrombits, int build) {
    if (!out || !root_index) return wfail(BBM_E_ARG, "bbm_wire_block_fname: bad argument");
    int n;
    if (minChrom != maxChrom) n = snprintf(out, cap, "%s%d/chr%d-%d_index_k%d_c%d_b%d.block", root_index, build, minChrom, maxChrom, k, chrombits, build);
    else n = snprintf(out, cap, "%s%d/chr%d_index_k%d_c%d_b%d.block", root_index, build, minChrom, k, chrombits, build);
    return (n < 0 || (size_t)n >= cap) ? wfail(BBM_E_CAPACITY, "bbm_wire_block_fname: name too long") : BBM_OK;
}

// Block.write (Block.java:74-112, compress=true, copyOnWrite=false): sites raw into fname, delta-coded starts gzipped into fname+"2.gz"
extern "C" int bbm_wire_write_block(const char* fname, const int32_t* sites, int64_t nsites, const int32_t* starts, int64_t nstarts) {
    if (!fname || nsites < 0 || nstarts < 2 || !starts || (nsites && !sites)) return wfail(BBM_E_ARG, "bbm_wire_write_block: bad argument");
    if (int e = bbm_wire_write_int_array(fname, sites, nsites)) return e;
    std::vector<int32_t> x((size_t)nstarts);
    x[0] = starts[0];
    for (int64_t i = 1; i < nstarts; ++i) x[i] = starts[i] - starts[i - 1];
    const std::string f2 = std::string(fname) + "2.gz";
    return bbm_wire_write_int_array(f2.c_str(), x.data(), nstarts);
}

// Block.read (Block.java:129-160): both arrays, prefix sum over starts
extern "C" int bbm_wire_read_block(const char* fname, int32_t** sites, int64_t* nsites, int32_t** starts, int64_t* nstarts) {
    if (!fname || !sites || !nsites || !starts || !nstarts) return wfail(BBM_E_ARG, "bbm_wire_read_block: bad argument");
    if (int e = bbm_wire_read_int_array(fname, sites, nsites)) return e;
    const std::string f2 = std::string(fname) + "2.gz";
    if (int e = bbm_wire_read_int_array(f2.c_str(), starts, nstarts)) { free(*sites); *sites = nullptr; return e; }
    int32_t* b = *starts; int32_t sum = *nstarts ? b[0] : 0;
    for (int64_t i = 1; i < *nstarts; ++i) { sum += b[i]; b[i] = sum; }
    const int64_t ns = *nstarts - 1;           // Block(int[],int[]) asserts numStarts is a power of two
    if (ns < 1 || (ns & (ns - 1)) != 0 || b[ns] != *nsites) {
        free(*sites); free(*starts); *sites = *starts = nullptr;
        return wfail(BBM_E_ARG, std::string(fname) + ": starts/sites are not a Block (numStarts not a power of two, or starts[numStarts] != sites.length)");
    }
    return BBM_OK;
}

// ReadWrite.write(ChromosomeArray ca, "chrN.chrom.gz") (FastaToChromArrays2.java:347-353)
extern "C" int bbm_wire_write_chrom(const char* path, int32_t chromosome, const int8_t* array, int32_t len, int32_t minIndex, int32_t maxIndex, int8_t strand) {
    if (!path || len < 0 || (len && !array)) return wfail(BBM_E_ARG, "bbm_wire_write_chrom: bad argument");
    Sink s; if (!s.open(path)) return wfail(BBM_E_ARG, std::string("cannot create ") + path);
    put_header(s);
    s.u8(TC_OBJECT); s.u8(TC_CLASSDESC); s.utf("dna.ChromosomeArray"); s.u64(SUID_CHROMOSOME_ARRAY); s.u8(SC_SERIALIZABLE); s.u16(5);
    s.u8('I'); s.utf("chromosome"); s.u8('I'); s.utf("maxIndex"); s.u8('I'); s.utf("minIndex"); s.u8('B'); s.utf("strand");
    s.u8('['); s.utf("array"); s.u8(TC_STRING); s.utf("[B");
    s.u8(TC_ENDBLOCKDATA); s.u8(TC_NULL);
    s.u32((uint32_t)chromosome); s.u32((uint32_t)maxIndex); s.u32((uint32_t)minIndex); s.u8((unsigned char)strand);
    put_array_desc(s, "[B", SUID_BYTE_ARRAY); s.u32((uint32_t)len); s.put(array, (size_t)len);
    return s.close() ? BBM_OK : wfail(BBM_E_ARG, std::string("write failed: ") + path);
}

// ChromosomeArray.read (ChromosomeArray.java:63-71)
extern "C" int bbm_wire_read_chrom(const char* path, int32_t* chromosome, int8_t** array, int32_t* len, int32_t* minIndex, int32_t* maxIndex, int8_t* strand) {
    if (!path || !array || !len) return wfail(BBM_E_ARG, "bbm_wire_read_chrom: bad argument");
    Source s; if (!s.open(path)) return wfail(BBM_E_ARG, std::string("cannot open ") + path);
    if (int e = get_header(s, path)) return e;
    if (s.u8() != TC_OBJECT) return wfail(BBM_E_ARG, std::string(path) + ": expected an object");
    ClassDesc d; if (int e = get_class_desc(s, d, path)) return e;
    if (d.name != "dna.ChromosomeArray" || d.suid != SUID_CHROMOSOME_ARRAY) return wfail(BBM_E_ARG, std::string(path) + ": class is " + d.name + ", expected dna.ChromosomeArray");
    int32_t chrom = 0, mx = -1, mn = 0; int8_t st = 0; bool haveArray = false; int8_t* a = nullptr; int64_t n = 0;
    for (const FieldDesc& f : d.fields) {       // values follow in descriptor order
        if (f.type == 'I') { const int32_t v = (int32_t)s.u32(); if (f.name == "chromosome") chrom = v; else if (f.name == "maxIndex") mx = v; else if (f.name == "minIndex") mn = v; }
        else if (f.type == 'B') { const int8_t v = (int8_t)s.u8(); if (f.name == "strand") st = v; }
        else if (f.type == '[' && f.name == "array" && f.sig == "[B") {
            if (int e = get_array_head(s, "[B", SUID_BYTE_ARRAY, &n, path)) return e;
            a = (int8_t*)malloc((size_t)std::max<int64_t>(n, 1));
            if (!a) return wfail(BBM_E_ARG, "out of memory");
            s.get(a, (size_t)n); haveArray = true;
        } else { free(a); return wfail(BBM_E_ARG, std::string(path) + ": unexpected field " + f.name); }
    }
    if (!s.ok || !haveArray) { free(a); return wfail(BBM_E_ARG, std::string(path) + ": truncated ChromosomeArray"); }
    *array = a; *len = (int32_t)n;
    if (chromosome) *chromosome = chrom; if (minIndex) *minIndex = mn; if (maxIndex) *maxIndex = mx; if (strand) *strand = st;
    return BBM_OK;
}

// summary.txt (FastaToChromArrays2.java:229-250): "#..." comment lines, then key<TAB>value
extern "C" int bbm_wire_write_summary(const char* path, const bbm_genome_summary* g) {
    if (!path || !g) return wfail(BBM_E_ARG, "bbm_wire_write_summary: bad argument");
    FILE* f = fopen(path, "w"); if (!f) return wfail(BBM_E_ARG, std::string("cannot create ") + path);
    fprintf(f, "#Summary\n#Version\t%d\nchroms\t%lld\nbases\t%lld\ndefined\t%lld\nundefined\t%lld\ncontigs\t%lld\nscaffolds\t%lld\ninterpad\t%lld\n", (int)g->version,
            (long long)g->chroms, (long long)g->bases, (long long)g->defined, (long long)(g->bases - g->defined), (long long)g->contigs, (long long)g->scaffolds, (long long)g->interpad);
    if (g->name[0]) fprintf(f, "name\t%s\n", g->name);
    return fclose(f) == 0 ? BBM_OK : wfail(BBM_E_ARG, std::string("write failed: ") + path);
}

// Data.setGenome's reader of the same file (dna/Data.java: "chroms", "bases", "defined", ... keys; unknown keys are ignored here)
extern "C" int bbm_wire_read_summary(const char* path, bbm_genome_summary* g) {
    if (!path || !g) return wfail(BBM_E_ARG, "bbm_wire_read_summary: bad argument");
    FILE* f = fopen(path, "r"); if (!f) return wfail(BBM_E_ARG, std::string("cannot open ") + path);
    memset(g, 0, sizeof(*g));
    char line[4096];
    while (fgets(line, sizeof line, f)) {
        char* tab = strchr(line, '\t'); if (!tab) continue;
        *tab = 0; char* val = tab + 1; val[strcspn(val, "\r\n")] = 0;
        const char* key = line;
        if (!strcmp(key, "#Version")) g->version = atoi(val);
        else if (key[0] == '#') continue;
        else if (!strcmp(key, "chroms")) g->chroms = atoll(val);
        else if (!strcmp(key, "bases")) g->bases = atoll(val);
        else if (!strcmp(key, "defined")) g->defined = atoll(val);
        else if (!strcmp(key, "undefined")) g->undefined = atoll(val);
        else if (!strcmp(key, "contigs")) g->contigs = atoll(val);
        else if (!strcmp(key, "scaffolds")) g->scaffolds = atoll(val);
        else if (!strcmp(key, "interpad")) g->interpad = atoll(val);
        else if (!strcmp(key, "name")) { strncpy(g->name, val, sizeof(g->name) - 1); }
    }
    fclose(f);
    return BBM_OK;
}

// ---- ReformatReads.breakReads (current/jgi/ReformatReads.java:1179-1219) as AbstractMapThread.run applies it (AbstractMapThread.java:441-443) ----
extern "C" int bbm_break_reads(const int8_t* bases, const int8_t* quality, const int64_t* read_off, int64_t nreads, const int8_t* names, const int64_t* name_off,
                               int32_t paired, int32_t max_len, int32_t min_len, int64_t* n_out, int64_t* bases_out_len, int64_t* names_out_len,
                               int8_t* out_bases, int8_t* out_quality, int64_t* out_read_off, int8_t* out_names, int64_t* out_name_off, int64_t* src,
                               int32_t* piece_start) {
    if (!bases || !read_off || !names || !name_off || nreads < 0 || !n_out || !bases_out_len || !names_out_len)
        return wfail(BBM_E_ARG, "bbm_break_reads: bad argument");
    if (max_len <= 0 && min_len <= 0) return wfail(BBM_E_ARG, "bbm_break_reads: min or max read length must be positive");
    if (max_len > 0 && max_len < min_len) return wfail(BBM_E_ARG, "bbm_break_reads: max read length must be at least min read length");
    const int64_t mn = min_len > 0 ? min_len : 0;
    const bool fill = out_bases != nullptr;
    if (fill && (!out_read_off || !out_names || !out_name_off || !src || !piece_start)) return wfail(BBM_E_ARG, "bbm_break_reads: output buffer missing");
    int64_t n = 0, nb = 0, nn = 0;
    auto emit = [&](int64_t r, int64_t start, int64_t stop, int num) {
        const int64_t nlen = name_off[r + 1] - name_off[r];
        char suffix[24]; int slen = 0;
        if (num > 0) slen = snprintf(suffix, sizeof suffix, "_%d", num);
        if (fill) {
            out_read_off[n] = nb; out_name_off[n] = nn; src[n] = r; piece_start[n] = (int32_t)start;
            memcpy(out_bases + nb, bases + read_off[r] + start, (size_t)(stop - start));
            if (out_quality && quality) memcpy(out_quality + nb, quality + read_off[r] + start, (size_t)(stop - start));
            memcpy(out_names + nn, names + name_off[r], (size_t)nlen);
            memcpy(out_names + nn + nlen, suffix, (size_t)slen);
        }
        n++; nb += stop - start; nn += nlen + slen;
    };
    for (int64_t r = 0; r < nreads; ++r) {
        const int64_t len = read_off[r + 1] - read_off[r];
        if (len < mn) continue;                                         // dropped
        if (max_len < 1 || len <= max_len) { emit(r, 0, len, 0); continue; }
        if (paired) return wfail(BBM_E_ARG, "bbm_break_reads: paired input is incompatible with breaking reads (a read is longer than max_len)");
        const int64_t limit = len - mn;
        int num = 1;
        for (int64_t start = 0, stop = max_len; start < limit; ++num, start += max_len, stop += max_len)
            emit(r, start, stop < len ? stop : len, num);
    }
    if (fill) { out_read_off[n] = nb; out_name_off[n] = nn; }
    *n_out = n; *bases_out_len = nb; *names_out_len = nn;
    return BBM_OK;
}
