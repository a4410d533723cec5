// ctucopy_b200 -- command-line host of the B200-native CtuCopy hot path.
//
// Keeps the `ctucopy` command-line surface of the reference for this path: the same
// options and -C config files (opts::parse, src/io/opts.cc:644-846), the same list format
// ("in out [spk] [vadout]" per line, src/io/batch.cc:349-356), the same decoders (raw s16
// with byte order, A-law, mu-law, canonical 44-byte WAVE: src/io/in.cc:434-611,
// src/io/amulaw.h:20-53) and byte-identical writers (HTK, pfile, Kaldi ark+scp, raw, wave,
// VAD '0'/'1' files: src/io/out.cc:61-781, src/io/pfile.cc:435-592, src/vad/vad.h:40-75).
// Everything between decode and write -- the per-frame loop of BATCH::process -- runs on
// the GPU through the C ABI of libctucopy_b200.so; there is no CPU implementation here.
//
// Errors: message on stderr, exit status 255 (src/main.cpp:31-60 returns -1).
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../include/ctucopy_b200.h"

namespace {

struct HostOpts {
    std::string list, in, out, config, format_in, format_out_arg, pfilename, arkfilename, filevad, vad_out;
    bool big_in = false, big_out = false, verbose = false, quiet = false, fb_printself = false;
    // multi-GPU extensions (not in the reference)
    int gpus = 1;                 // -gpus N    : N worker threads in this process, one handle per GPU
    int shard_r = 0, shard_n = 1; // -shard r/N : this process handles shard r of N (one process per GPU)
    int merge_n = 0;              // -merge N   : merge the pfile / ark+scp shards of N finished `-shard` runs
    bool ss_carry = false;        // -ss_carry on: hwss / fwss / 2fwss with the reference's list semantics (a file's noise estimate starts
                                  // from the enhanced last frame of the file before it, src/nr/nr.cc:212-222); one GPU, list order
    int device = -1;              // -device d  : CUDA device (default: shard index modulo device count)
    std::string stat_cmvn, apply_cmvn;   // CMVN statistics file to write / to apply (src/io/opts.cc:678-685)
    std::string ark_ref;          // ark path recorded in scp lines (the merged file's name)
};

struct ListEntry { std::string in, out, spk, vadout; };

[[noreturn]] void die(const std::string &m) { throw std::runtime_error(m); }

uint32_t bswap32(uint32_t x) { return ((x & 0xFFu) << 24) | ((x & 0xFF00u) << 8) | ((x & 0xFF0000u) >> 8) | ((x & 0xFF000000u) >> 24); }
uint16_t bswap16(uint16_t x) { return (uint16_t)((x >> 8) | (x << 8)); }

std::vector<unsigned char> slurp(const std::string &path, const char *err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) die(err);
    std::vector<unsigned char> b;
    unsigned char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) b.insert(b.end(), buf, buf + n);
    std::fclose(f);
    return b;
}

// "<ark path>" -> "<scp path>" the way the reference derives it: split on '.', drop the token
// "ark" and everything after it, append "scp" (src/io/out.cc:756-774; empty tokens vanish)
std::string ark_to_scp(const std::string &ark) {
    std::string line, tok;
    std::vector<std::string> toks;
    for (char ch : ark) { if (ch == '.') { if (!tok.empty()) toks.push_back(tok); tok.clear(); } else tok += ch; }
    if (!tok.empty()) toks.push_back(tok);
    for (auto &t : toks) { if (t == "ark") return line + "scp"; line += t + "."; }
    return line + "scp";
}

std::string pfile_header(uint64_t nsent, uint64_t frames, uint64_t dim);

struct Writers {
    const HostOpts &o;
    const ctu_config &c;
    int dim;
    FILE *pf = nullptr, *ark = nullptr, *scp = nullptr;
    bool rows_formatted = false;   // the rows arrive in container layout (ctu_plan_set_row_format): big-endian floats / pfile rows
    std::vector<uint32_t> sent_table{0};
    uint64_t pf_frames = 0;
    Writers(const HostOpts &o_, const ctu_config &c_, int dim_) : o(o_), c(c_), dim(dim_) {
        std::string fo(c.format_out);
        if (fo == "pfile") {
            pf = std::fopen(o.pfilename.c_str(), "wb");
            if (!pf) die("ERROR: Can not open the pfile: " + o.pfilename + ".");
            std::vector<char> z(32768, 0);
            std::fwrite(z.data(), 1, z.size(), pf);
        } else if (fo == "ark") {
            ark = std::fopen(o.arkfilename.c_str(), "wb");
            if (!ark) die("OUT: Cannot create output ark file!");
            scp = std::fopen(ark_to_scp(o.arkfilename).c_str(), "wt");
            if (!scp) die("Cannot open output scp file for writing!");
        }
    }
    // htkOUT::new_file, src/io/out.cc:145-158.  Two reference quirks are part of the header: the "T bit" is the DECIMAL
    // constant 100000 cut to 16 bits (:158), and with -fea_trap on sample input the first written row renames the kind
    // to "spec" (:182), so every file after the first of the list carries base kind 8.
    int parmkind(size_t list_index) const {
        std::string k(c.fea_kind);
        if (c.fea_trap && !c.fea_in && list_index > 0) k = "spec";
        int kind = k == "lpc" ? 11 : k == "dctc" ? 6 : k == "trapdct" ? 9 : k == "spec" ? 8 : k == "logspec" ? 7 : 9;
        std::string k0(c.fea_kind);
        bool c0 = c.fea_c0 && !(k0 == "lpa" || k0 == "spec" || k0 == "logspec");
        if (c0) kind |= 020000;
        if (c.fea_E) kind |= 000100;
        if (c.fea_delta && c.n_order >= 1) kind |= 000400;
        if (c.fea_delta && c.n_order >= 2) kind |= 001000;
        if (c.fea_delta && c.n_order == 3) kind = (kind | 100000) & 0xffff;
        return kind;
    }
    void features(const ListEntry &e, const float *rows, int64_t n, size_t list_index) {
        std::string fo(c.format_out);
        if (fo == "htk") {
            FILE *f = std::fopen(e.out.c_str(), "wb");
            if (!f) die("OUT: Cannot create output file!");
            uint32_t frames = (uint32_t)n, period = (uint32_t)std::floor(.5 + 10000000. * c.wshift / (double)c.fs);
            uint16_t size = (uint16_t)(4 * dim), kind = (uint16_t)parmkind(list_index);
            if (o.big_out) { frames = bswap32(frames); period = bswap32(period); size = bswap16(size); kind = bswap16(kind); }
            std::fwrite(&frames, 4, 1, f); std::fwrite(&period, 4, 1, f); std::fwrite(&size, 2, 1, f); std::fwrite(&kind, 2, 1, f);
            if (o.big_out && rows_formatted) {
                std::fwrite(rows, 4, (size_t)n * dim, f);             // byte-swapped on the device
            } else if (o.big_out) {
                std::vector<uint32_t> t((size_t)n * dim);
                std::memcpy(t.data(), rows, t.size() * 4);
                for (auto &v : t) v = bswap32(v);
                std::fwrite(t.data(), 4, t.size(), f);
            } else {
                std::fwrite(rows, 4, (size_t)n * dim, f);
            }
            std::fclose(f);
        } else if (fo == "pfile") {
            // rows: u32 sentence, u32 frame, float32 x dim, all big-endian (src/io/pfile.cc:470-539)
            std::vector<uint32_t> row(dim + 2);
            uint32_t sid = (uint32_t)sent_table.size() - 1;
            if (rows_formatted) std::fwrite(rows, 4 * (size_t)(dim + 2), (size_t)n, pf);      // whole rows from the device
            else for (int64_t t = 0; t < n; t++) {
                row[0] = bswap32(sid); row[1] = bswap32((uint32_t)t);
                std::memcpy(row.data() + 2, rows + t * dim, (size_t)dim * 4);
                for (int i = 0; i < dim; i++) row[2 + i] = bswap32(row[2 + i]);
                std::fwrite(row.data(), 4, row.size(), pf);
            }
            pf_frames += (uint64_t)n;
            sent_table.push_back((uint32_t)pf_frames);
        } else if (fo == "ark") {
            // "<key> \0BFM \4<rows>\4<cols><data>"; the scp points at the \0 (src/io/out.cc:680-700)
            std::fprintf(ark, "%s %cBFM %c", e.out.c_str(), 0, 4);
            uint32_t r = (uint32_t)n; int32_t cols = dim;
            std::fwrite(&r, 4, 1, ark);
            std::fprintf(ark, "%c", 4);
            std::fwrite(&cols, 4, 1, ark);
            long long idx = (long long)std::ftell(ark) - 15;
            std::fprintf(scp, "%s %s:%llu\n", e.out.c_str(), (o.ark_ref.empty() ? o.arkfilename : o.ark_ref).c_str(), (unsigned long long)idx);
            std::fwrite(rows, 4, (size_t)n * dim, ark);
        }
    }
    void waveform(const ListEntry &e, const int16_t *s, int64_t n) {
        FILE *f = std::fopen(e.out.c_str(), "wb");
        if (!f) die("OUT: Cannot open output stream!");
        std::string fo(c.format_out);
        if (fo == "wave") {
            uint32_t data = (uint32_t)(2 * n), whole = data + 36, fs = (uint32_t)c.fs, bps = fs * 2, pcm = 16;
            uint16_t one = 1, two = 2, sixteen = 16;
            std::fwrite("RIFF", 1, 4, f); std::fwrite(&whole, 4, 1, f); std::fwrite("WAVE", 1, 4, f); std::fwrite("fmt ", 1, 4, f);
            std::fwrite(&pcm, 4, 1, f); std::fwrite(&one, 2, 1, f); std::fwrite(&one, 2, 1, f); std::fwrite(&fs, 4, 1, f);
            std::fwrite(&bps, 4, 1, f); std::fwrite(&two, 2, 1, f); std::fwrite(&sixteen, 2, 1, f);
            std::fwrite("data", 1, 4, f); std::fwrite(&data, 4, 1, f);
            std::fwrite(s, 2, (size_t)n, f);
        } else if (o.big_out) {
            std::vector<uint16_t> t((size_t)n);
            for (int64_t i = 0; i < n; i++) t[i] = bswap16((uint16_t)s[i]);
            std::fwrite(t.data(), 2, t.size(), f);
        } else {
            std::fwrite(s, 2, (size_t)n, f);
        }
        std::fclose(f);
    }
    void vad(const ListEntry &e, const uint8_t *v, int64_t n) {
        FILE *f = std::fopen(e.vadout.c_str(), "wb");
        if (!f) die("FileWriter: cannot open file!");
        for (int64_t i = 0; i < n; i++) std::fputc(v[i] ? '1' : '0', f);
        std::fclose(f);
    }
    // -vad_out_mode debug: the side files of FileWriter(filename, suffix) (src/vad/vad.h:39-76) next to the decision file.
    // steps: CTU_VAD_DEBUG_COLS doubles per VAD step (criterion, threshold, three state values), vad0: unfiltered decisions.
    // VAD::save_frame (src/vad/vad.cc:710-725) runs after the state has advanced and only once the majority filter has a
    // decision: row i holds step min(i + (order-1)/2, T-1); the *init flags compare the incremented frame index.
    void vad_debug(const ListEntry &e, const double *steps, const uint8_t *vad0, int64_t T) {
        const int64_t h = (c.vad_filter_order - 1) / 2;
        auto open = [&](const char *suffix) {
            FILE *f = std::fopen((e.vadout + "_" + suffix).c_str(), "wb");
            if (!f) die("FileWriter: cannot open file!");
            return f;
        };
        auto step = [&](int64_t i) { return std::min(i + h, T - 1); };
        const int64_t rows = (T > h) ? T : 0;
        auto doubles = [&](const char *suffix, int col) {
            FILE *f = open(suffix);
            std::vector<double> v((size_t)rows);
            for (int64_t i = 0; i < rows; i++) v[(size_t)i] = steps[step(i) * CTU_VAD_DEBUG_COLS + col];
            std::fwrite(v.data(), 8, v.size(), f);
            std::fclose(f);
        };
        auto constant = [&](const char *suffix, double x) {
            FILE *f = open(suffix);
            std::vector<double> v((size_t)rows, x);
            std::fwrite(v.data(), 8, v.size(), f);
            std::fclose(f);
        };
        auto flags = [&](const char *suffix, int init) {       // '1' while the (incremented) frame index is <= init
            FILE *f = open(suffix);
            for (int64_t i = 0; i < rows; i++) std::fputc(step(i) + 1 <= init ? '1' : '0', f);
            std::fclose(f);
        };
        { FILE *f = open("vad0"); for (int64_t i = 0; i < rows; i++) std::fputc(vad0[step(i)] ? '1' : '0', f); std::fclose(f); }
        if (!std::strcmp(c.vad_cri_mode, "energy")) doubles("energy", 0);
        else { doubles("cepdist", 0); flags("c0init", c.vad_cepdist_init); }
        const std::string m(c.vad_thr_mode);
        if (m == "absolute") constant("thr", c.vad_absolute_thr);
        else if (m == "perc") { doubles("crimin", 2); doubles("crimax", 3); doubles("thr", 1); }
        else if (m == "adapt") { flags("init", c.vad_adapt_init); doubles("crimean", 2); doubles("crimean2", 3); doubles("crivar", 4); doubles("thr", 1); }
        else { doubles("dmin", 2); doubles("dmax", 3); doubles("dyn", 4); constant("dynmin", c.vad_dyn_min); doubles("thr", 1); }
    }
    void close() {
        if (pf) {
            // sentence table right behind the data, then the ASCII header (src/io/pfile.cc:435-468, 573-592)
            uint64_t nsent = sent_table.size() - 1;
            for (uint32_t v : sent_table) { uint32_t b = bswap32(v); std::fwrite(&b, 4, 1, pf); }
            std::string h = pfile_header(nsent, pf_frames, (uint64_t)dim);
            std::fseek(pf, 0, SEEK_SET);
            std::fwrite(h.data(), 1, h.size(), pf);
            std::fclose(pf); pf = nullptr;
        }
        if (ark) { std::fclose(ark); ark = nullptr; }
        if (scp) { std::fclose(scp); scp = nullptr; }
    }
};

void take_host_option(HostOpts &o, const char *l, const char *r) {
    std::string opt(l);
    auto val = [&]() { return r ? std::string(r) : std::string(); };
    if (opt == "-S" && r) o.list = r;
    else if (opt == "-i" && r) o.in = r;
    else if (opt == "-o" && r) o.out = r;
    else if (opt == "-format_in" && r) o.format_in = r;
    else if (opt == "-format_out" && r) {
        std::string v = val();
        size_t eq = v.find('=');
        if (v.find("pfile=") != std::string::npos) o.pfilename = v.substr(eq + 1);
        else if (v.find("ark=") != std::string::npos) o.arkfilename = v.substr(eq + 1);
    } else if (opt == "-endian_in" && r) { if (val() == "big") o.big_in = true; else if (val() == "little") o.big_in = false; }
    else if (opt == "-endian_out" && r) { if (val() == "big") o.big_out = true; else if (val() == "little") o.big_out = false; }
    else if (opt == "-vad" && r) { std::string v = val(); if (v.find("file=") != std::string::npos) o.filevad = v.substr(v.find('=') + 1); }
    else if (opt == "-vad_out" && r) o.vad_out = r;
    else if (opt == "-stat_cmvn" && r) o.stat_cmvn = r;
    else if (opt == "-apply_cmvn" && r) o.apply_cmvn = r;
    else if (opt == "-v" || opt == "-verbose") { o.verbose = true; o.quiet = false; }
    else if (opt == "-quiet") { o.quiet = true; o.verbose = false; }
    else if (opt == "-fb_printself") o.fb_printself = true;
    else if (opt == "-gpus" && r) o.gpus = std::max(1, std::atoi(r));
    else if (opt == "-device" && r) o.device = std::atoi(r);
    else if (opt == "-merge" && r) o.merge_n = std::atoi(r);
    else if (opt == "-ss_carry") o.ss_carry = !r || val() == "on";
    else if (opt == "-shard" && r) {
        if (std::sscanf(r, "%d/%d", &o.shard_r, &o.shard_n) != 2 || o.shard_n < 1 || o.shard_r < 0 || o.shard_r >= o.shard_n)
            die("OPTS: -shard takes r/N with 0 <= r < N");
    }
}

bool is_host_extension(const char *opt) {
    return !std::strcmp(opt, "-gpus") || !std::strcmp(opt, "-device") || !std::strcmp(opt, "-merge") || !std::strcmp(opt, "-shard") ||
           !std::strcmp(opt, "-ss_carry");
}

// samples a file will decode to, from its size / header only (used to balance shards and to
// place each shard inside a list-wide external VAD file)
// HTK parameter file header (htkIN::new_file, src/io/in.cc:630-650): only the vector size is used
int htk_width(FILE *f, bool big) {
    unsigned char b[12];
    std::fseek(f, 0, SEEK_SET);
    if (std::fread(b, 1, 12, f) != 12) return -1;
    const int bytes = big ? ((b[8] << 8) | b[9]) : (b[8] | (b[9] << 8));
    return bytes / 4;
}

int64_t count_samples(const HostOpts &o, const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) die(o.format_in == "wave" ? "IN: Cannot open file!" : "IN: Cannot open data file!");
    std::fseek(f, 0, SEEK_END);
    int64_t size = (int64_t)std::ftell(f), n = 0;
    if (o.format_in == "htk") {                  // rows: the reference reads whole vectors until the file ends
        const int wd = htk_width(f, o.big_in);
        if (wd <= 0) { std::fclose(f); die("IN: Cannot read the HTK header of " + path); }
        n = (size - 12) / ((int64_t)wd * 4);
    } else
    if (o.format_in == "raw") n = size / 2;
    else if (o.format_in == "alaw" || o.format_in == "mulaw") n = size;
    else if (o.format_in == "wave") {
        unsigned char b[44] = {0};
        std::fseek(f, 0, SEEK_SET);
        if (std::fread(b, 1, 44, f) == 44) {
            int64_t d = (int64_t)((uint32_t)b[40] | ((uint32_t)b[41] << 8) | ((uint32_t)b[42] << 16) | ((uint32_t)b[43] << 24));
            n = std::min<int64_t>(d / 2, (size - 44) / 2);
        }
    }
    std::fclose(f);
    return n;
}

// "<dir>/shard<r>of<n>_<file>": the prefix keeps the ".ark" token where ark_to_scp expects it
std::string shard_name(const std::string &base, int r, int n) {
    size_t sl = base.rfind('/');
    std::string dir = sl == std::string::npos ? "" : base.substr(0, sl + 1), file = sl == std::string::npos ? base : base.substr(sl + 1);
    return dir + "shard" + std::to_string(r) + "of" + std::to_string(n) + "_" + file;
}

// contiguous ranges of list lines, balanced by cumulative sample count
std::vector<size_t> partition(const std::vector<int64_t> &samples, int n) {
    std::vector<size_t> cut(n + 1, samples.size());
    cut[0] = 0;
    long double total = 0, acc = 0;
    for (int64_t v : samples) total += (long double)v;
    int r = 1;
    for (size_t i = 0; i < samples.size() && r < n; i++) {
        acc += (long double)samples[i];
        while (r < n && acc >= total * r / n) cut[r++] = i + 1;
    }
    return cut;
}

std::vector<unsigned char> slurp_or_empty(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) die("MERGE: missing shard file " + path);
    std::fclose(f);
    return slurp(path, "MERGE: cannot read shard");
}

// Kaldi ark + scp: byte concatenation, scp offsets shifted by the bytes that precede the shard
void merge_ark(const std::string &ark, int n) {
    const std::string scp = ark_to_scp(ark);
    FILE *fa = std::fopen(ark.c_str(), "wb"), *fs = std::fopen(scp.c_str(), "wt");
    if (!fa || !fs) die("MERGE: cannot create " + ark);
    unsigned long long base = 0;
    for (int r = 0; r < n; r++) {
        auto b = slurp_or_empty(shard_name(ark, r, n));
        std::ifstream sl(ark_to_scp(shard_name(ark, r, n)));
        std::string line;
        while (std::getline(sl, line)) {
            size_t c = line.rfind(':');
            if (c == std::string::npos) continue;
            unsigned long long off = std::strtoull(line.c_str() + c + 1, nullptr, 10);
            std::fprintf(fs, "%s:%llu\n", line.substr(0, c).c_str(), off + base);
        }
        std::fwrite(b.data(), 1, b.size(), fa);
        base += b.size();
        std::remove(shard_name(ark, r, n).c_str());
        std::remove(ark_to_scp(shard_name(ark, r, n)).c_str());
    }
    std::fclose(fa); std::fclose(fs);
}

std::string pfile_header(uint64_t nsent, uint64_t frames, uint64_t dim) {
    // ASCII header of an ICSI pfile as the reference writes it (src/io/pfile.cc:435-468)
    auto num = [](uint64_t v) { return std::to_string((unsigned long long)v); };
    const uint64_t ncol = dim + 2;
    std::string h;
    h += "-pfile_header version 0 size 32768\n";
    h += "-num_sentences " + num(nsent) + "\n";
    h += "-num_frames " + num(frames) + "\n";
    h += "-first_feature_column 2\n";
    h += "-num_features " + num(dim) + "\n";
    h += "-first_label_column " + num(dim + 2) + "\n";
    h += "-num_labels 0\n";
    h += "-format dd" + std::string((size_t)dim, 'f') + "\n";
    h += "-data size " + num(ncol * frames) + " offset 0 ndim 2 nrow " + num(frames) + " ncol " + num(ncol) + "\n";
    h += "-sent_table_data size " + num(nsent + 1) + " offset " + num(ncol * frames) + " ndim 1\n";
    h += "-end\n";
    return h;
}

// pfile: rows concatenated with the sentence id rewritten, sentence table rebuilt
void merge_pfile(const std::string &pf, int n) {
    FILE *fo = std::fopen(pf.c_str(), "wb");
    if (!fo) die("MERGE: cannot create " + pf);
    std::vector<char> z(32768, 0);
    std::fwrite(z.data(), 1, z.size(), fo);
    std::vector<uint32_t> table{0};
    uint64_t frames = 0, dim = 0;
    for (int r = 0; r < n; r++) {
        auto b = slurp_or_empty(shard_name(pf, r, n));
        if (b.size() < 32768) die("MERGE: short pfile shard");
        std::string head((const char *)b.data(), 32768);
        auto field = [&](const char *key) -> uint64_t {
            size_t p = head.find(key);
            if (p == std::string::npos) die(std::string("MERGE: pfile shard lacks ") + key);
            return std::strtoull(head.c_str() + p + std::strlen(key), nullptr, 10);
        };
        const uint64_t ns = field("-num_sentences "), nf = field("-num_frames "), d = field("-num_features ");
        if (r && d != dim) die("MERGE: pfile shards differ in feature count");
        dim = d;
        const size_t ncol = (size_t)dim + 2;
        if (b.size() < 32768 + 4 * (ncol * nf + ns + 1)) die("MERGE: truncated pfile shard");
        uint32_t *rows = reinterpret_cast<uint32_t *>(b.data() + 32768);
        const uint32_t sent0 = (uint32_t)table.size() - 1;
        for (uint64_t t = 0; t < nf; t++) rows[t * ncol] = bswap32(bswap32(rows[t * ncol]) + sent0);
        std::fwrite(rows, 4, ncol * nf, fo);
        const uint32_t *st = rows + ncol * nf;
        for (uint64_t k = 1; k <= ns; k++) table.push_back((uint32_t)frames + bswap32(st[k]));
        frames += nf;
        std::remove(shard_name(pf, r, n).c_str());
    }
    for (uint32_t v : table) { uint32_t x = bswap32(v); std::fwrite(&x, 4, 1, fo); }
    std::string h = pfile_header(table.size() - 1, frames, dim);
    std::fseek(fo, 0, SEEK_SET);
    std::fwrite(h.data(), 1, h.size(), fo);
    std::fclose(fo);
}

// ---- streaming pipeline ----------------------------------------------------------------------
// The list range is cut into batches of ~64 Mi samples (128 MB of PCM: pinning memory costs ~0.4 ms per MB).  Three stages run concurrently on three
// batch slots: (1) read + decode the batch's files straight into page-locked memory (several
// threads; every file's place in the batch is known from its size before it is read),
// (2) the GPU (ctu_run: chunked H2D / kernels / D2H), (3) the writers (per-utterance files by
// several threads, containers in list order).  Replaces the reference's file loop
// (src/io/batch.cc:349-412) end to end.
// reader and writer threads of a batch (each stage has its own pool); CTU_IO_THREADS overrides
const int IO_THREADS = std::getenv("CTU_IO_THREADS") ? std::max(1, std::atoi(std::getenv("CTU_IO_THREADS"))) : 8;

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
const bool g_timing = std::getenv("CTU_TIMING") != nullptr;     // stage times on stderr
double g_t0 = now_s();
void tmark(const char *what, double t0) { if (g_timing) std::fprintf(stderr, "[ctu timing] %-28s %.4f s   (ends at %.3f)\n", what, now_s() - t0, now_s() - g_t0); }

void parallel_for(size_t n, int nthreads, const std::function<void(size_t)> &fn) {
    std::atomic<size_t> next{0};
    std::vector<std::string> errs((size_t)nthreads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            try { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); }
            catch (const std::exception &e) { errs[t] = e.what(); if (errs[t].empty()) errs[t] = "unknown error"; next = n; }
        });
    for (auto &x : th) x.join();
    for (auto &e : errs) if (!e.empty()) die(e);
}

// decode one file into dst[0 .. n) (n = count_samples of the same file)
void decode_into(const HostOpts &o, int fs, const std::string &path, int16_t *dst, int64_t n) {
    const std::string &fmt = o.format_in;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) die(fmt == "wave" ? "IN: Cannot open file!" : "IN: Cannot open data file!");
    auto rd = [&](void *p, size_t bytes) { if (bytes && std::fread(p, 1, bytes, f) != bytes) { std::fclose(f); die("IN: Error reading input file!"); } };
    if (fmt == "raw") {
        rd(dst, (size_t)n * 2);
        if (o.big_in) for (int64_t i = 0; i < n; i++) dst[i] = (int16_t)bswap16((uint16_t)dst[i]);
    } else if (fmt == "alaw" || fmt == "mulaw") {
        // (the batch pipeline ships the codes to the GPU as they are, ctu_plan_run_host_g711; this host-side expansion
        // serves the CMVN passes)
        std::vector<unsigned char> b((size_t)n);
        rd(b.data(), (size_t)n);
        int16_t table[256];
        ctu_g711_table(fmt == "alaw", table);
        for (int64_t i = 0; i < n; i++) dst[i] = table[b[(size_t)i]];
    } else if (fmt == "wave") {
        unsigned char b[44];
        if (std::fread(b, 1, 44, f) != 44 || std::memcmp(b, "RIFF", 4)) { std::fclose(f); die("IN: No RIFF header in file!"); }
        if (std::memcmp(b + 8, "WAVE", 4)) { std::fclose(f); die("IN: Not a WAVE file!"); }
        auto u16 = [&](size_t p) { return (int)(b[p] | (b[p + 1] << 8)); };
        auto u32 = [&](size_t p) { return (uint32_t)b[p] | ((uint32_t)b[p + 1] << 8) | ((uint32_t)b[p + 2] << 16) | ((uint32_t)b[p + 3] << 24); };
        if (u16(20) != 1) { std::fclose(f); die("IN: Not a PCM WAVE file!"); }
        if ((long)u32(24) != fs) { std::fclose(f); die("IN: WAVE file reports different sampling rate than specified!"); }
        if (u16(22) != 1) { std::fclose(f); die("IN: Input WAVE file is not mono!"); }
        if (u16(34) != 16) { std::fclose(f); die("IN: Not 16 bits per sample!"); }
        rd(dst, (size_t)n * 2);
    } else {
        std::fclose(f);
        die("IN: Unknown input file format!");
    }
    std::fclose(f);
}

// one HTK parameter file into dst[n rows x in_dim]: the first `width` floats of a row come from the file, the rest is 0.
// The library takes rows of -nfeacoefs floats; the delta chain reads the first fea_ncepcoefs+1 of them.
void read_features_into(const HostOpts &o, const ctu_config &cfg, const std::string &path, float *dst, int64_t n) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) die("IN: Cannot open data file!");
    const int wd = htk_width(f, o.big_in), in_dim = cfg.nfeacoefs;
    const bool chain = cfg.fea_delta && cfg.n_order > 0;
    if (wd > in_dim) { std::fclose(f); die("IN: " + path + " has more columns than -nfeacoefs (the reference overruns its vector)"); }
    if (chain ? wd < cfg.fea_ncepcoefs + 1 : wd != in_dim) {
        std::fclose(f);
        die("IN: " + path + (chain ? " has fewer than fea_ncepcoefs+1 columns" : " does not have -nfeacoefs columns"));
    }
    std::vector<uint32_t> row((size_t)wd);
    for (int64_t t = 0; t < n; t++) {
        if (std::fread(row.data(), 4, (size_t)wd, f) != (size_t)wd) { std::fclose(f); die("IN: Error reading input file!"); }
        if (o.big_in) for (auto &v : row) v = bswap32(v);
        std::memcpy(dst + t * in_dim, row.data(), (size_t)wd * 4);
        for (int i = wd; i < in_dim; i++) dst[t * in_dim + i] = 0.f;
    }
    std::fclose(f);
}

// a file's bytes as they are (G.711 codes: expanded on the GPU)
void read_bytes_into(const std::string &path, unsigned char *dst, int64_t n) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) die("IN: Cannot open data file!");
    if (n && std::fread(dst, 1, (size_t)n, f) != (size_t)n) { std::fclose(f); die("IN: Error reading input file!"); }
    std::fclose(f);
}

struct Pinned {
    void *p = nullptr; uint64_t cap = 0;
    void reserve(uint64_t bytes) {
        if (bytes <= cap) return;
        ctu_host_free(p); p = nullptr; cap = 0;
        if (ctu_host_alloc(&p, bytes)) die(ctu_last_error(nullptr));
        cap = bytes;
    }
    ~Pinned() { ctu_host_free(p); }
};

struct Batch {
    size_t i0 = 0, i1 = 0;                 // list range
    std::vector<int64_t> off, frames, rows;
    int64_t total = 0, total_os = 0;       // frames, output samples
    size_t ext_pos = 0;
    Pinned pcm, fea, wav, vout, vnr;
    std::vector<double> vdbg;              // -vad_out_mode debug: per-step records and unfiltered decisions
    std::vector<uint8_t> vad0;
    int state = 0;                         // 0 free, 1 read, 2 computed
};

void process_range(const HostOpts &ho, const ctu_config &cfg, const std::vector<ListEntry> &list, size_t r0, size_t r1, int device,
                   const std::vector<unsigned char> &extvad, size_t frame0, uint64_t rand0 = 0) {
    const bool do_vad = std::strcmp(cfg.vad_apply_mode, "none") || std::strcmp(cfg.vad_out_mode, "none");
    const bool vad_file = std::strcmp(cfg.vad_out_mode, "none") != 0;
    const bool vad_dbg = !std::strcmp(cfg.vad_out_mode, "debug");
    const bool use_ext = !std::strcmp(cfg.vadmode, "file");
    const double t_start = now_s();
    ctu_handle *h = nullptr;
    if (ctu_create(&cfg, device, &h)) die(ctu_last_error(nullptr));
    tmark("ctu_create (CUDA init)", t_start);
    if (ho.ss_carry && ctu_set_option(h, "ss_carry", 1)) die(ctu_last_error(h));
    ctu_set_rand_offset(h, rand0);            // -dither: rand() values the list lines before this range have drawn
    const int dim = ctu_feature_dim(h);
    const bool sig = ctu_is_signal_output(h);
    const int in_dim = ctu_input_dim(h);          // > 0: the list names feature files (-format_in htk), offsets count rows
    const bool fea_in = in_dim > 0;
    const bool g711 = (ho.format_in == "alaw" || ho.format_in == "mulaw");   // 8-bit codes go to the GPU as they are
    const std::string fo(cfg.format_out);
    const bool per_file = (fo == "htk" || sig);
    Writers W(ho, cfg, dim);
    // pfile rows and byte-swapped HTK rows are laid out by the device; the writers then copy whole rows
    const int row_fmt = (sig || fea_in) ? CTU_ROWS_NATIVE : (fo == "pfile") ? CTU_ROWS_PFILE : (fo == "htk" && ho.big_out) ? CTU_ROWS_BE : CTU_ROWS_NATIVE;
    const int row_words = dim + (row_fmt == CTU_ROWS_PFILE ? 2 : 0);
    W.rows_formatted = row_fmt != CTU_ROWS_NATIVE;
    // sizes first: batch boundaries and every file's offset are known before anything is read
    const size_t nfiles = r1 - r0;
    std::vector<int64_t> nsamp(nfiles);
    double t1 = now_s();
    parallel_for(nfiles, IO_THREADS, [&](size_t i) { nsamp[i] = count_samples(ho, list[r0 + i].in); });
    tmark("file sizes", t1);
    std::vector<size_t> cuts{r0};
    {
        int64_t acc = 0;
        for (size_t i = 0; i < nfiles; i++) {
            acc += fea_in ? nsamp[i] * in_dim * 2 : nsamp[i];     // the same bytes per batch either way
            if (acc >= (int64_t(1) << 26) || (int64_t)(r0 + i + 1 - cuts.back()) >= (1 << 20)) { cuts.push_back(r0 + i + 1); acc = 0; }
        }
        if (cuts.back() != r1) cuts.push_back(r1);
    }
    const size_t nb = cuts.size() - 1;
    Batch slots[3];
    std::mutex mu;
    std::condition_variable cv;
    std::string err;
    auto fail = [&](const std::string &m) { std::lock_guard<std::mutex> l(mu); if (err.empty()) err = m.empty() ? "unknown error" : m; cv.notify_all(); };

    // ---- stage 1: reader ------------------------------------------------------------------------
    std::thread reader([&]() {
        try {
            size_t fr_before = frame0;
            for (size_t k = 0; k < nb; k++) {
                Batch &B = slots[k % 3];
                { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return B.state == 0 || !err.empty(); }); if (!err.empty()) return; }
                B.i0 = cuts[k]; B.i1 = cuts[k + 1];
                const size_t n = B.i1 - B.i0;
                B.off.assign(n + 1, 0); B.frames.assign(n, 0); B.rows.assign(n, 0);
                B.total = 0; B.total_os = 0;
                for (size_t u = 0; u < n; u++) {
                    const int64_t N = nsamp[B.i0 - r0 + u];
                    B.off[u + 1] = B.off[u] + N;
                    const int64_t T = ctu_num_frames(h, N);
                    if (T < 0) die("IO: Signal shorter than one frame!");
                    B.total += T;
                    B.total_os += ctu_num_output_samples(h, N);
                }
                B.ext_pos = fr_before;
                fr_before += (size_t)B.total;
                B.pcm.reserve(fea_in ? (uint64_t)(B.off[n] + 1) * in_dim * 4 : (uint64_t)(B.off[n] + 8) * 2);
                int16_t *pcm = (int16_t *)B.pcm.p;
                float *fin = (float *)B.pcm.p;
                double tr = now_s();
                parallel_for(n, IO_THREADS, [&](size_t u) {
                    if (fea_in) read_features_into(ho, cfg, list[B.i0 + u].in, fin + B.off[u] * in_dim, B.off[u + 1] - B.off[u]);
                    else if (g711) read_bytes_into(list[B.i0 + u].in, (unsigned char *)B.pcm.p + B.off[u], B.off[u + 1] - B.off[u]);
                    else decode_into(ho, cfg.fs, list[B.i0 + u].in, pcm + B.off[u], B.off[u + 1] - B.off[u]);
                });
                tmark("batch read+decode", tr);
                { std::lock_guard<std::mutex> l(mu); B.state = 1; }
                cv.notify_all();
            }
        } catch (const std::exception &e) { fail(e.what()); }
    });
    // ---- stage 3: writer ------------------------------------------------------------------------
    std::thread writer([&]() {
        try {
            for (size_t k = 0; k < nb; k++) {
                Batch &B = slots[k % 3];
                { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return B.state == 2 || !err.empty(); }); if (!err.empty()) return; }
                const size_t n = B.i1 - B.i0;
                std::vector<int64_t> row0(n + 1, 0), s0(n + 1, 0);
                for (size_t u = 0; u < n; u++) { row0[u + 1] = row0[u] + B.frames[u]; s0[u + 1] = s0[u] + ctu_num_output_samples(h, B.off[u + 1] - B.off[u]); }
                const float *fea = (const float *)B.fea.p;
                const int16_t *wav = (const int16_t *)B.wav.p;
                const uint8_t *vout = (const uint8_t *)B.vout.p;
                auto one = [&](size_t u) {
                    const ListEntry &e = list[B.i0 + u];
                    if (sig) W.waveform(e, wav + s0[u], s0[u + 1] - s0[u]);
                    else W.features(e, fea + row0[u] * row_words, B.rows[u], B.i0 + u);
                    if (do_vad && vad_file) W.vad(e, vout + row0[u], B.frames[u]);
                    if (vad_dbg) W.vad_debug(e, B.vdbg.data() + row0[u] * CTU_VAD_DEBUG_COLS, B.vad0.data() + row0[u], B.frames[u]);
                };
                double tw = now_s();
                if (per_file) parallel_for(n, IO_THREADS, one);      // one file per utterance: any order
                else for (size_t u = 0; u < n; u++) one(u);          // pfile / ark: list order
                tmark("batch write", tw);
                if (ho.verbose)
                    for (size_t u = 0; u < n; u++) std::cerr << "processing: " << list[B.i0 + u].in << " - " << B.frames[u] << " frames." << std::endl;
                { std::lock_guard<std::mutex> l(mu); B.state = 0; }
                cv.notify_all();
            }
        } catch (const std::exception &e) { fail(e.what()); }
    });
    // ---- stage 2: GPU (this thread) ----------------------------------------------------------------
    try {
        for (size_t k = 0; k < nb; k++) {
            Batch &B = slots[k % 3];
            { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return B.state == 1 || !err.empty(); }); if (!err.empty()) break; }
            const int n = (int)(B.i1 - B.i0);
            if (!sig) B.fea.reserve((uint64_t)B.total * row_words * 4);
            if (sig) B.wav.reserve((uint64_t)B.total_os * 2);
            B.vout.reserve((uint64_t)B.total + 1); B.vnr.reserve((uint64_t)B.total + 1);
            const uint8_t *ev = nullptr;
            if (use_ext) {
                if (B.ext_pos + (size_t)B.total > extvad.size()) die("NR: Unexpected end of VAD file!");
                ev = extvad.data() + B.ext_pos;
            }
            double tg = now_s();
            if (fea_in || g711 || vad_dbg || row_fmt != CTU_ROWS_NATIVE) {
                ctu_plan *pl = nullptr;
                if (ctu_plan_create(h, B.off.data(), n, &pl)) die(ctu_last_error(h));
                if (row_fmt != CTU_ROWS_NATIVE && ctu_plan_set_row_format(pl, row_fmt, (uint32_t)(B.i0 - r0))) { ctu_plan_destroy(pl); die(ctu_last_error(h)); }
                const int st = fea_in ? ctu_plan_run_host_fea(pl, (const float *)B.pcm.p, (float *)B.fea.p)
                               : g711 ? ctu_plan_run_host_g711(pl, (const uint8_t *)B.pcm.p, ho.format_in == "alaw", ev, sig ? nullptr : (float *)B.fea.p,
                                                               sig ? (int16_t *)B.wav.p : nullptr, (uint8_t *)B.vnr.p, (uint8_t *)B.vout.p)
                                      : ctu_plan_run_host(pl, (const int16_t *)B.pcm.p, ev, sig ? nullptr : (float *)B.fea.p, sig ? (int16_t *)B.wav.p : nullptr,
                                                          (uint8_t *)B.vnr.p, (uint8_t *)B.vout.p);
                if (st) { ctu_plan_destroy(pl); die(ctu_last_error(h)); }
                ctu_plan_frames_per_utt(pl, B.frames.data());
                ctu_plan_rows_per_utt(pl, B.rows.data());
                if (vad_dbg) {
                    B.vdbg.resize((size_t)B.total * CTU_VAD_DEBUG_COLS); B.vad0.resize((size_t)B.total);
                    if (ctu_plan_fetch_vad_debug(pl, B.vdbg.data(), B.vad0.data())) { ctu_plan_destroy(pl); die(ctu_last_error(h)); }
                }
                ctu_plan_destroy(pl);
            } else
            if (ctu_run(h, (const int16_t *)B.pcm.p, B.off.data(), n, ev, sig ? nullptr : (float *)B.fea.p, B.total, sig ? (int16_t *)B.wav.p : nullptr,
                        B.total_os, (uint8_t *)B.vnr.p, (uint8_t *)B.vout.p, B.frames.data(), B.rows.data()))
                die(ctu_last_error(h));
            tmark("batch ctu_run", tg);
            { std::lock_guard<std::mutex> l(mu); B.state = 2; }
            cv.notify_all();
        }
    } catch (const std::exception &e) { fail(e.what()); }
    reader.join();
    writer.join();
    if (!err.empty()) { ctu_destroy(h); die(err); }
    W.close();
    ctu_destroy(h);
    tmark("whole range", t_start);
}

// ---- CMVN over the whole list (src/io/batch.cc:136-152, 339-420; src/fea/post_impl.cc:52-118) ------------
// -stat_cmvn <f>: two passes, per-speaker mean and variance of every feature, written to <f>, no features.
// -apply_cmvn <f> with <f> not there yet: the same two passes, <f> written, then a third pass that writes
// (F - mean) / var.  (With an existing <f> the reference never finds its speakers again and writes +-inf;
// that mode is refused.)  The device does the column sums and the normalisation; the host groups the
// utterances by speaker in list order and owns the file.
void process_cmvn(const HostOpts &ho, const ctu_config &cfg, const std::vector<ListEntry> &list, int device,
                  const std::vector<unsigned char> &extvad, int ngpus = 1) {
    const bool apply = !ho.apply_cmvn.empty();
    const std::string statfile = apply ? ho.apply_cmvn : ho.stat_cmvn;
    if (apply) {
        if (FILE *f = std::fopen(statfile.c_str(), "rb")) {
            std::fclose(f);
            die("CTU: -apply_cmvn with an existing statistics file: the reference does not find its speakers again and writes +-inf "
                "(src/io/in.cc:743-770, src/fea/post_impl.cc:45); delete the file to compute and apply the statistics in one run");
        }
        std::cout << "IN: Cannot open stat. cmvn file!" << std::endl << "IN: Stat. cmvn file is being created: " << statfile << std::endl;
    }
    // speakers: third list column (second one for statistics-only lists of two columns)
    std::vector<std::string> names;
    std::vector<int> spk(list.size());
    for (size_t i = 0; i < list.size(); i++) {
        std::string id = list[i].spk;
        if (id.empty()) {
            if (apply) die(" Bad list format for applying cmvn!");
            id = list[i].out;
        }
        size_t j = 0;
        while (j < names.size() && names[j] != id) j++;
        if (j == names.size()) names.push_back(id);
        spk[i] = (int)j;
    }
    // -gpus N: one handle and one worker thread per GPU; the batches of a pass are dealt to the workers.  What a batch
    // contributes is kept PER UTTERANCE and added up on the host in list order after the pass, so the statistics (and the
    // text of the statistics file) are the single-process ones bit for bit, whatever the number of GPUs or their timing:
    // the one cross-GPU reduction of the path is a host-side sum of n_utts x dim doubles.
    const int G = std::max(1, ngpus);
    const std::string fo(cfg.format_out);
    if (G > 1 && apply && fo != "htk") die("CTU: -apply_cmvn under -gpus N writes one HTK file per utterance; pfile / ark containers need one process");
    std::vector<ctu_handle *> hs((size_t)G, nullptr);
    for (int g = 0; g < G; g++)
        if (ctu_create(&cfg, G > 1 ? g % std::max(1, ctu_device_count()) : device, &hs[(size_t)g])) die(ctu_last_error(nullptr));
    ctu_handle *h = hs[0];
    const int fdim = ctu_feature_dim(h), dim = ctu_cmvn_dim(h);
    const int in_dim = ctu_input_dim(h);          // > 0: feature files in (-format_in htk)
    const bool fea_in = in_dim > 0;
    const bool use_ext = !std::strcmp(cfg.vadmode, "file");
    std::vector<int64_t> nsamp(list.size());
    parallel_for(list.size(), IO_THREADS, [&](size_t i) { nsamp[i] = count_samples(ho, list[i].in); });
    std::vector<size_t> cuts{0};
    // samples per batch (64 Mi by default; CTU_CMVN_BATCH_SAMPLES lets a test cut a short list into several batches)
    const int64_t batch_samples = std::getenv("CTU_CMVN_BATCH_SAMPLES") ? std::max<long long>(1, std::atoll(std::getenv("CTU_CMVN_BATCH_SAMPLES"))) : (int64_t(1) << 26);
    {
        int64_t acc = 0;
        for (size_t i = 0; i < list.size(); i++) {
            acc += fea_in ? nsamp[i] * in_dim * 2 : nsamp[i];
            if (acc >= batch_samples) { cuts.push_back(i + 1); acc = 0; }
        }
        if (cuts.back() != list.size()) cuts.push_back(list.size());
    }
    const size_t nbatch = cuts.size() - 1;
    // frames of the batches before each batch (external VAD bytes are one list-wide stream)
    std::vector<int64_t> nfr(list.size());
    for (size_t i = 0; i < list.size(); i++) { nfr[i] = ctu_num_frames(h, nsamp[i]); if (nfr[i] < 0) die("IO: Signal shorter than one frame!"); }
    std::vector<size_t> fr_before(nbatch + 1, 0);
    for (size_t k = 0; k < nbatch; k++) { size_t t = 0; for (size_t i = cuts[k]; i < cuts[k + 1]; i++) t += (size_t)nfr[i]; fr_before[k + 1] = fr_before[k] + t; }
    const size_t ns = names.size();
    std::vector<double> sum(ns * dim, 0.0), var(ns * dim, 0.0), cnt(ns, 0.0);
    std::vector<double> per_utt(list.size() * (size_t)dim, 0.0);       // what each utterance adds in the current pass
    Writers W(ho, cfg, fdim);
    for (int pass = 0; pass < (apply ? 3 : 2); pass++) {
        std::atomic<size_t> next{0};
        std::vector<std::string> errs((size_t)G);
        auto worker = [&](int g) {
            try {
                ctu_handle *hg = hs[(size_t)g];
                Pinned pcmbuf, feabuf;
                for (size_t k; (k = next.fetch_add(1)) < nbatch;) {
                    const size_t i0 = cuts[k], i1 = cuts[k + 1], n = i1 - i0;
                    std::vector<int64_t> off(n + 1, 0), frames(n);
                    int64_t total = 0;
                    for (size_t u = 0; u < n; u++) { off[u + 1] = off[u] + nsamp[i0 + u]; frames[u] = nfr[i0 + u]; total += frames[u]; }
                    pcmbuf.reserve(fea_in ? (uint64_t)(off[n] + 1) * in_dim * 4 : (uint64_t)(off[n] + 8) * 2);
                    int16_t *pcm = (int16_t *)pcmbuf.p;
                    float *fin = (float *)pcmbuf.p;
                    parallel_for(n, std::max(1, IO_THREADS / G), [&](size_t u) {
                        if (fea_in) read_features_into(ho, cfg, list[i0 + u].in, fin + off[u] * in_dim, off[u + 1] - off[u]);
                        else decode_into(ho, cfg.fs, list[i0 + u].in, pcm + off[u], off[u + 1] - off[u]);
                    });
                    const uint8_t *ev = nullptr;
                    if (use_ext) {
                        if (fr_before[k] + (size_t)total > extvad.size()) die("NR: Unexpected end of VAD file!");
                        ev = extvad.data() + fr_before[k];
                    }
                    ctu_plan *p = nullptr;
                    if (ctu_plan_create(hg, off.data(), (int)n, &p)) die(ctu_last_error(hg));
                    if (fea_in ? ctu_plan_run_host_fea(p, fin, nullptr) : ctu_plan_run_host_keep(p, pcm, ev)) die(ctu_last_error(hg));
                    std::vector<double> a(n * dim), b(n * dim);
                    if (pass == 0) {
                        if (ctu_plan_colsums(p, nullptr, per_utt.data() + i0 * dim)) die(ctu_last_error(hg));
                    } else if (pass == 1) {
                        for (size_t u = 0; u < n; u++) std::copy(sum.begin() + (size_t)spk[i0 + u] * dim, sum.begin() + (size_t)(spk[i0 + u] + 1) * dim, a.begin() + u * dim);
                        if (ctu_plan_colsums(p, a.data(), per_utt.data() + i0 * dim)) die(ctu_last_error(hg));
                    } else {
                        for (size_t u = 0; u < n; u++) {
                            std::copy(sum.begin() + (size_t)spk[i0 + u] * dim, sum.begin() + (size_t)(spk[i0 + u] + 1) * dim, a.begin() + u * dim);
                            std::copy(var.begin() + (size_t)spk[i0 + u] * dim, var.begin() + (size_t)(spk[i0 + u] + 1) * dim, b.begin() + u * dim);
                        }
                        if (ctu_plan_normalise(p, a.data(), b.data())) die(ctu_last_error(hg));
                        feabuf.reserve((uint64_t)total * fdim * 4);
                        if (ctu_plan_fetch(p, (float *)feabuf.p, nullptr, nullptr, nullptr)) die(ctu_last_error(hg));
                        int64_t r0 = 0;
                        for (size_t u = 0; u < n; u++) { W.features(list[i0 + u], (const float *)feabuf.p + r0 * fdim, frames[u], i0 + u); r0 += frames[u]; }
                    }
                    ctu_plan_destroy(p);
                }
            } catch (const std::exception &e) { errs[(size_t)g] = e.what(); if (errs[(size_t)g].empty()) errs[(size_t)g] = "unknown error"; next = nbatch; }
        };
        if (G == 1) worker(0);
        else {
            std::vector<std::thread> th;
            for (int g = 0; g < G; g++) th.emplace_back(worker, g);
            for (auto &t : th) t.join();
        }
        for (auto &e : errs) if (!e.empty()) die(e);
        // the pass's sums, utterance by utterance in list order (cmvn_POST::sum_fea / sum_cv, src/fea/post_impl.cc:52-97)
        if (pass == 0) {
            for (size_t i = 0; i < list.size(); i++) {
                const int j = spk[i];
                for (int c = 0; c < dim; c++) sum[(size_t)j * dim + c] += per_utt[i * dim + c];
                cnt[j] += (double)nfr[i];
            }
            for (size_t j = 0; j < ns; j++) for (int c = 0; c < dim; c++) sum[j * dim + c] /= cnt[j];       // sum -> mean (stat_cm)
        } else if (pass == 1) {
            for (size_t i = 0; i < list.size(); i++)
                for (int c = 0; c < dim; c++) var[(size_t)spk[i] * dim + c] += per_utt[i * dim + c];
            for (size_t j = 0; j < ns; j++) for (int c = 0; c < dim; c++) var[j * dim + c] /= (cnt[j] - 1);  // stat_cv
            // statistics file (cmvnOUT::save_frame, src/io/out.cc:591-615) in the reference's order: the
            // internal vector F[1..], then F[0] -- c0 of the static block goes last, everything else stays.  Stacked rows
            // (-fea_trap) are written as they are, so the vector the statistics see is the written one; with feature-file
            // input nothing is skipped or rotated at all (src/fea/post_impl.cc:56-58)
            const std::string kind(cfg.fea_kind);
            const bool cep = (kind == "dctc" || kind == "lpc") && !cfg.fea_trap;
            const int nblk = cfg.fea_ncepcoefs + 1;
            std::vector<int> wcol(dim);            // statistics index -> writer column
            for (int w = 0; w < dim; w++) {
                int i = w;                         // internal index of writer column w
                if (cep) { const int j = w / nblk, pw = w % nblk; i = j * nblk + (pw == nblk - 1 ? 0 : pw + 1); }
                wcol[i == 0 ? dim - 1 : i - 1] = w;
            }
            if (fea_in) for (int w = 0; w < dim; w++) wcol[w] = w;
            FILE *f = std::fopen(statfile.c_str(), "wt");
            if (!f) die("OUT: Cannot create output file with stat. of cmvn!");
            for (size_t j = 0; j < ns; j++) {
                std::fprintf(f, "%s\nmean\t", names[j].c_str());
                for (int i = 0; i < dim; i++) std::fprintf(f, i + 1 < dim ? "%f " : "%f", sum[j * dim + wcol[i]]);
                std::fprintf(f, "\nvar\t");
                for (int i = 0; i < dim; i++) std::fprintf(f, i + 1 < dim ? "%f " : "%f\n", var[j * dim + wcol[i]]);
            }
            std::fclose(f);
        }
    }
    W.close();
    for (auto *x : hs) ctu_destroy(x);
}

int run(int argc, char **argv) {
    if (argc == 1) { std::cerr << "usage: ctucopy_b200 <ctucopy options> -S <list> [-gpus N | -shard r/N | -merge N]   (options: man/ctucopy4 of the reference)" << std::endl; die("OPTS: No command line options!"); }
    ctu_config cfg;
    ctu_config_init(&cfg);
    HostOpts ho;
    // host-owned options: config file first, then the command line (same order as the library's parser)
    for (int j = 1; j + 1 < argc; j++) {
        if (std::strcmp(argv[j], "-C")) continue;
        std::ifstream cf(argv[j + 1]);
        std::string line;
        while (std::getline(cf, line)) {
            size_t h = line.find('#');
            if (h != std::string::npos) line.resize(h);
            char a[1024] = "", b[1024] = "";
            int n = std::sscanf(line.c_str(), "%1023s %1023s", a, b);
            if (n >= 1) take_host_option(ho, a, n >= 2 ? b : nullptr);
        }
    }
    std::vector<const char *> lib_argv;         // the library's parser does not know the multi-GPU extensions
    for (int j = 1; j < argc; j++) {
        if (argv[j][0] != '-') { lib_argv.push_back(argv[j]); continue; }
        const char *r = (j + 1 < argc && argv[j + 1][0] != '-') ? argv[j + 1] : nullptr;
        take_host_option(ho, argv[j], r);
        if (is_host_extension(argv[j])) { if (r) j++; continue; }
        lib_argv.push_back(argv[j]);
    }
    if (ctu_config_parse(&cfg, (int)lib_argv.size(), lib_argv.data())) die(ctu_config_error());
    if (ho.merge_n > 0) {                        // merge step of a multi-process run: no GPU involved
        if (!ho.pfilename.empty()) merge_pfile(ho.pfilename, ho.merge_n);
        if (!ho.arkfilename.empty()) merge_ark(ho.arkfilename, ho.merge_n);
        return 0;
    }
    if (ho.fb_printself) {
        int32_t nb = 0;
        if (ctu_design_filter_bank(&cfg, nullptr, nullptr, nullptr, &nb)) die(ctu_last_error(nullptr));
        std::vector<double> mat((size_t)nb * cfg.wfftby2);
        std::vector<int32_t> lo(nb), hi(nb);
        ctu_design_filter_bank(&cfg, mat.data(), lo.data(), hi.data(), &nb);
        for (int b = 0; b < nb; b++) { for (int i = 0; i < cfg.wfftby2; i++) std::cerr << mat[(size_t)b * cfg.wfftby2 + i] << "\t"; std::cerr << std::endl; }
    }
    // the list (single-file mode is one list line)
    std::vector<ListEntry> list;
    const bool do_vad = std::strcmp(cfg.vad_apply_mode, "none") || std::strcmp(cfg.vad_out_mode, "none");
    if (!ho.in.empty() || !ho.out.empty()) {
        if (ho.in.empty() || ho.out.empty()) die("OPTS: Single file mode has to be set at both sides (input and output)!");
        list.push_back({ho.in, ho.out, "", ho.vad_out});
    } else {
        if (ho.list.empty()) die("BATCH: Nothing to do!");
        std::ifstream lf(ho.list);
        if (!lf) die("BATCH: Cannot open list file!");
        std::string line;
        while (std::getline(lf, line)) {
            std::vector<std::string> tok;
            size_t i = 0;
            while (i < line.size()) {
                while (i < line.size() && (line[i] == ' ' || line[i] == '\t')) i++;
                size_t b = i;
                while (i < line.size() && line[i] != ' ' && line[i] != '\t') i++;
                if (i > b) tok.push_back(line.substr(b, i - b));
            }
            if (tok.size() < 2) die("BATCH: Bad list format!");
            if (do_vad && tok.size() < 4) die("BATCH: Bad list format!");
            list.push_back({tok[0], tok[1], tok.size() > 2 ? tok[2] : "", tok.size() > 3 ? tok[3] : ""});
        }
    }
    std::vector<unsigned char> extvad;
    if (!std::strcmp(cfg.vadmode, "file")) extvad = slurp(ho.filevad, "NR: Unable to open VAD file!\n");
    const int nparts = ho.gpus > 1 ? ho.gpus : ho.shard_n;
    if (ho.ss_carry && (nparts != 1 || !ho.stat_cmvn.empty() || !ho.apply_cmvn.empty()))
        die("CTU: -ss_carry chains every file to the one before it: one GPU, one process, no CMVN passes");
    if (!ho.stat_cmvn.empty() || !ho.apply_cmvn.empty()) {
        if (ho.shard_n != 1) die("CTU: CMVN statistics span the whole list: run it in one process (-gpus N uses N GPUs from one process; -shard cannot)");
        process_cmvn(ho, cfg, list, ho.device < 0 ? 0 : ho.device, extvad, ho.gpus);
        return 0;
    }
    if (nparts == 1) {
        process_range(ho, cfg, list, 0, list.size(), ho.device < 0 ? 0 : ho.device, extvad, 0);
        return 0;
    }
    // ---- utterance sharding: contiguous list ranges balanced by sample count, no exchange between shards
    std::vector<int64_t> samples(list.size());
    std::vector<size_t> frames_before(list.size() + 1, 0);
    std::vector<uint64_t> drawn_before(list.size() + 1, 0);      // -dither: one rand() per loaded sample (src/io/in.cc:452-455)
    for (size_t i = 0; i < list.size(); i++) {
        samples[i] = count_samples(ho, list[i].in);
        int64_t T = samples[i] < cfg.window - cfg.wshift ? 0 : (samples[i] - (cfg.window - cfg.wshift)) / cfg.wshift;
        frames_before[i + 1] = frames_before[i] + (size_t)T;
        drawn_before[i + 1] = drawn_before[i] + (uint64_t)T * cfg.wshift + (cfg.window - cfg.wshift);
    }
    const std::vector<size_t> cut = partition(samples, nparts);
    auto shard_opts = [&](int r) {
        HostOpts o = ho;
        if (!o.arkfilename.empty()) { o.ark_ref = ho.arkfilename; o.arkfilename = shard_name(ho.arkfilename, r, nparts); }
        if (!o.pfilename.empty()) o.pfilename = shard_name(ho.pfilename, r, nparts);
        return o;
    };
    if (ho.gpus <= 1) {                          // one process per GPU (e.g. under torchrun): this is shard r
        process_range(shard_opts(ho.shard_r), cfg, list, cut[ho.shard_r], cut[ho.shard_r + 1], ho.device < 0 ? ho.shard_r : ho.device, extvad,
                      frames_before[cut[ho.shard_r]], drawn_before[cut[ho.shard_r]]);
        return 0;
    }
    std::vector<std::string> errs(nparts);
    std::vector<std::thread> th;
    for (int r = 0; r < nparts; r++)
        th.emplace_back([&, r]() {
            try { process_range(shard_opts(r), cfg, list, cut[r], cut[r + 1], r % std::max(1, ctu_device_count()), extvad, frames_before[cut[r]], drawn_before[cut[r]]); }
            catch (const std::exception &e) { errs[r] = e.what(); if (errs[r].empty()) errs[r] = "unknown error"; }
        });
    for (auto &t : th) t.join();
    for (auto &e : errs) if (!e.empty()) die(e);
    if (!ho.pfilename.empty()) merge_pfile(ho.pfilename, nparts);
    if (!ho.arkfilename.empty()) merge_ark(ho.arkfilename, nparts);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    try {
        const int rc = run(argc, argv);
        // every output file is closed by now: skip the CUDA context teardown (about a second on a B200 host)
        std::fflush(nullptr);
        std::_Exit(rc);
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return -1;
    }
}
