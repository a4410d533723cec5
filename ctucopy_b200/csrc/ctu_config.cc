// Host-side option handling for the hot path: defaults, presets, one-option setter,
// argv walk and derived sizes.  Behaviour follows `class opts` of the reference
// (src/io/opts.cc:27-325, 644-846) so that the same option strings configure the same
// pipeline; the implementation is table driven.
#include "ctu_internal.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

static thread_local std::string g_cfg_err;

const char *ctu_config_error(void) { return g_cfg_err.c_str(); }

static int cfg_fail(const std::string &m) {
    g_cfg_err = m;
    return CTU_ERR_CONFIG;
}

static void put(char *dst, const char *src, size_t cap = CTU_STR) {
    std::snprintf(dst, cap, "%s", src);
}

int ctu_config_init(ctu_config *c) {
    if (!c) return CTU_ERR_CONFIG;
    std::memset(c, 0, sizeof(*c));
    c->abi_version = CTU_ABI_VERSION;
    put(c->format_out, "");
    c->fs = 0;
    c->preem = 0.0f;
    c->dither = 0.0;  // code default is 0 although the man page says 1 (src/io/opts.cc:38)
    c->remove_dc = 1;
    c->remove_dc1 = 0;
    c->window_ms = 25.0;
    c->wshift_ms = 10.0;
    put(c->fb_scale, "mel");
    put(c->fb_shape, "triang");
    c->fb_norm = c->fb_power = c->fb_eqld = c->fb_inld = 1;
    put(c->fb_definition, "26filters", CTU_FBDEF);
    put(c->vadmode, "none");
    put(c->nr_mode, "none");
    c->nr_p = 0.95; c->nr_q = 0.99; c->nr_a = 1.0; c->nr_b = 1.0;
    c->nr_initsegs = 10;
    c->nr_when = 0;
    put(c->fea_kind, "lpc");
    c->fea_lporder = 12; c->fea_ncepcoefs = 12; c->fea_c0 = 1; c->fea_E = 0; c->fea_rawenergy = 0;
    c->fea_Z_exp = -1.f; c->fea_Z_block = -1.f; c->cms_exp_coef = -1.f; c->stat_cmvn = 0; c->apply_cmvn = 0;   // src/io/opts.cc:86-91
    c->fea_lifter = 22;
    c->fea_trapdct_traplen = 0; c->fea_trapdct_ndct = 0;  // the reference leaves these uninitialised
    c->fea_delta = 0; c->n_order = 0; c->d_win = c->a_win = c->t_win = 2;
    c->fea_trap = 0; c->trap_win = 5; c->fea_in = 0; c->nfeacoefs = 13;      // src/io/opts.cc:97-99
    put(c->filters, "", CTU_FBDEF); c->weight_of_td_iir_mfcc_bank = 2.026f;  // src/io/opts.cc:100, 104
    put(c->vad_apply_mode, "none");
    put(c->vad_out_mode, "none");
    put(c->vad_cri_mode, "energy");
    put(c->vad_thr_mode, "perc");
    c->vad_energy_db = 1;
    put(c->vad_cepdist_mode, "lpc");
    c->vad_cepdist_p = 0.8; c->vad_cepdist_init = 4; c->vad_lpc_coefs = 14;
    c->vad_absolute_thr = 1.0;
    c->vad_perc_init = 10; c->vad_perc_thr = 50.0;
    c->vad_adapt_init = 20; c->vad_adapt_q = 0.9; c->vad_adapt_za = 2.0;
    c->vad_dyn_init = 5; c->vad_dyn_perc = 50.0; c->vad_dyn_min = 1.0;
    c->vad_dyn_qmaxinc = 0.8; c->vad_dyn_qmaxdec = 0.995; c->vad_dyn_qmindec = 0.8; c->vad_dyn_qmininc = 0.9999;
    c->vad_filter_order = 3;
    return CTU_OK;
}

// presets are macros applied at parse position (src/io/opts.cc:196-253)
static int apply_preset(ctu_config *c, const char *name) {
    std::string p(name);
    if (p == "mfcc") {
        put(c->fb_scale, "mel"); put(c->fb_shape, "triang"); c->fb_power = 1;
        put(c->fb_definition, "1-26/26filters", CTU_FBDEF);
        put(c->nr_mode, "none");
        c->fb_eqld = 0; c->fb_inld = 0;
        put(c->fea_kind, "dctc"); c->fea_ncepcoefs = 12; c->fea_c0 = 1; c->fea_E = 0;
        c->fea_lifter = 22; c->fea_rawenergy = 0;
    } else if (p == "plpc") {
        put(c->fb_scale, "bark"); put(c->fb_shape, "trapez"); c->fb_power = 1;
        put(c->fb_definition, "1-15/15filters", CTU_FBDEF);
        put(c->nr_mode, "none");
        c->fb_eqld = 1; c->fb_inld = 1;
        put(c->fea_kind, "lpc"); c->fea_lporder = 12; c->fea_ncepcoefs = 12; c->fea_c0 = 1; c->fea_E = 0;
        c->fea_lifter = 22; c->fea_rawenergy = 0;
    } else if (p == "exten") {
        c->window_ms = 32.0; c->wshift_ms = 16.0;
        put(c->fb_definition, "none", CTU_FBDEF); put(c->fb_scale, "none"); put(c->fb_shape, "none");
        c->nr_a = 2.0;
        c->fb_eqld = c->fb_inld = c->fb_power = c->fb_norm = 0;
        put(c->nr_mode, "exten"); put(c->fea_kind, "none");
        c->fea_c0 = 0; c->fea_E = 0; c->fea_lifter = 0; c->fea_rawenergy = 0;
    } else {
        return cfg_fail("OPTS: Unknown preset!");
    }
    return CTU_OK;
}

namespace {
enum Kind { K_INT, K_DBL, K_FLT, K_ONOFF, K_STR, K_IGNORE, K_FLAG_IGNORE };
struct Opt { const char *name; Kind kind; size_t off; };
#define O(field) offsetof(ctu_config, field)
const Opt kOpts[] = {
    {"-fs", K_INT, O(fs)}, {"-preem", K_FLT, O(preem)}, {"-dither", K_DBL, O(dither)},
    {"-remove_dc", K_ONOFF, O(remove_dc)}, {"-remove_dc1", K_ONOFF, O(remove_dc1)},
    {"-w", K_DBL, O(window_ms)}, {"-s", K_DBL, O(wshift_ms)},
    {"-fb_scale", K_STR, O(fb_scale)}, {"-fb_shape", K_STR, O(fb_shape)},
    {"-fb_norm", K_ONOFF, O(fb_norm)}, {"-fb_power", K_ONOFF, O(fb_power)},
    {"-fb_eqld", K_ONOFF, O(fb_eqld)}, {"-fb_inld", K_ONOFF, O(fb_inld)},
    {"-nr_mode", K_STR, O(nr_mode)}, {"-nr_p", K_DBL, O(nr_p)}, {"-nr_q", K_DBL, O(nr_q)},
    {"-nr_a", K_DBL, O(nr_a)}, {"-nr_b", K_DBL, O(nr_b)}, {"-nr_initsegs", K_INT, O(nr_initsegs)},
    {"-d_win", K_INT, O(d_win)}, {"-a_win", K_INT, O(a_win)}, {"-t_win", K_INT, O(t_win)},
    {"-fea_lporder", K_INT, O(fea_lporder)}, {"-fea_ncepcoefs", K_INT, O(fea_ncepcoefs)},
    {"-fea_c0", K_ONOFF, O(fea_c0)}, {"-fea_E", K_ONOFF, O(fea_E)}, {"-fea_rawenergy", K_ONOFF, O(fea_rawenergy)},
    {"-fea_lifter", K_INT, O(fea_lifter)},
    {"-vad_apply_mode", K_STR, O(vad_apply_mode)}, {"-vad_out_mode", K_STR, O(vad_out_mode)},
    {"-vad_cri_mode", K_STR, O(vad_cri_mode)}, {"-vad_thr_mode", K_STR, O(vad_thr_mode)},
    {"-vad_energy_db", K_ONOFF, O(vad_energy_db)}, {"-vad_cepdist_mode", K_STR, O(vad_cepdist_mode)},
    {"-vad_cepdist_p", K_DBL, O(vad_cepdist_p)}, {"-vad_cepdist_init", K_INT, O(vad_cepdist_init)},
    {"-vad_lpc_coefs", K_INT, O(vad_lpc_coefs)}, {"-vad_absolute_thr", K_DBL, O(vad_absolute_thr)},
    {"-vad_perc_init", K_INT, O(vad_perc_init)}, {"-vad_perc_thr", K_DBL, O(vad_perc_thr)},
    {"-vad_adapt_init", K_INT, O(vad_adapt_init)}, {"-vad_adapt_q", K_DBL, O(vad_adapt_q)},
    {"-vad_adapt_za", K_DBL, O(vad_adapt_za)}, {"-vad_dyn_init", K_INT, O(vad_dyn_init)},
    {"-vad_dyn_perc", K_DBL, O(vad_dyn_perc)}, {"-vad_dyn_min", K_DBL, O(vad_dyn_min)},
    {"-vad_dyn_qmaxinc", K_DBL, O(vad_dyn_qmaxinc)}, {"-vad_dyn_qmaxdec", K_DBL, O(vad_dyn_qmaxdec)},
    {"-vad_dyn_qmindec", K_DBL, O(vad_dyn_qmindec)}, {"-vad_dyn_qmininc", K_DBL, O(vad_dyn_qmininc)},
    {"-vad_filter_order", K_INT, O(vad_filter_order)},
    {"-fea_Z_exp", K_FLT, O(fea_Z_exp)}, {"-fea_Z_block", K_FLT, O(fea_Z_block)},
    {"-weight_of_td_iir_mfcc_bank", K_FLT, O(weight_of_td_iir_mfcc_bank)},
    // owned by the host CLI, not by the hot path: accepted, no effect here
    {"-S", K_IGNORE, 0}, {"-i", K_IGNORE, 0}, {"-o", K_IGNORE, 0}, {"-C", K_IGNORE, 0},
    {"-endian_in", K_IGNORE, 0}, {"-endian_out", K_IGNORE, 0},
    {"-vad_out", K_IGNORE, 0}, {"-nr_rasta", K_IGNORE, 0},
    {"-online_in", K_FLAG_IGNORE, 0}, {"-online_out", K_FLAG_IGNORE, 0}, {"-fb_printself", K_FLAG_IGNORE, 0},
    {"-verbose", K_FLAG_IGNORE, 0}, {"-v", K_FLAG_IGNORE, 0}, {"-quiet", K_FLAG_IGNORE, 0},
    {"-info", K_FLAG_IGNORE, 0}, {"-h", K_FLAG_IGNORE, 0}, {"--help", K_FLAG_IGNORE, 0},
};
#undef O
}  // namespace

int ctu_config_set(ctu_config *c, const char *l, const char *r) {
    if (!c || !l) return cfg_fail("OPTS: null option");
    std::string opt(l);
    char *base = reinterpret_cast<char *>(c);
    // options with structured values first
    if (opt == "-format_out") {
        if (!r) return CTU_OK;
        std::string v(r);
        if (v.find("pfile=") != std::string::npos) put(c->format_out, "pfile");
        else if (v.find("ark=") != std::string::npos) put(c->format_out, "ark");
        else put(c->format_out, r);
        return CTU_OK;
    }
    if (opt == "-fb_definition") { if (r) put(c->fb_definition, r, CTU_FBDEF); return CTU_OK; }
    if (opt == "-preset") { return r ? apply_preset(c, r) : CTU_OK; }
    if (opt == "-vad") {
        if (!r) return CTU_OK;
        std::string v(r);
        if (v == "burg") put(c->vadmode, "burg");
        else if (v.find("file=") != std::string::npos) put(c->vadmode, "file");
        else return cfg_fail("OPTS: Syntax error in option -vad !");
        return CTU_OK;
    }
    if (opt == "-stat_cmvn") { if (r) c->stat_cmvn = 1; return CTU_OK; }
    if (opt == "-apply_cmvn") { if (r) c->apply_cmvn = 1; return CTU_OK; }
    if (opt == "-nr_when") {
        if (r && !std::strcmp(r, "beforeFB")) c->nr_when = 0;
        else if (r && !std::strcmp(r, "afterFB")) c->nr_when = 1;
        return CTU_OK;
    }
    if (opt == "-fea_kind") {
        if (!r) return cfg_fail("OPTS: Missing argument to '-fea_kind' option!");
        std::string v(r);
        if (v.find("trapdct") != std::string::npos) {
            size_t c1 = v.find(',');
            if (c1 == std::string::npos)
                return cfg_fail("OPTS: Syntax error in option -fea_kind! (should be -fea_kind trapdct,<X>,<Y>)");
            size_t c2 = v.find(',', c1 + 1);
            if (c2 == std::string::npos)
                return cfg_fail("OPTS: Syntax error in option -fea_kind! (should be -fea_kind trapdct,<X>,<Y>)");
            put(c->fea_kind, v.substr(0, c1).c_str());
            c->fea_trapdct_traplen = std::atoi(v.c_str() + c1 + 1);
            c->fea_trapdct_ndct = std::atoi(v.c_str() + c2 + 1);
        } else {
            put(c->fea_kind, r);
        }
        return CTU_OK;
    }
    if (opt == "-fea_delta") {
        if (!r) return CTU_OK;
        c->fea_delta = 1;
        c->fea_trap = 0;
        if (!std::strcmp(r, "d")) c->n_order = 1;
        else if (!std::strcmp(r, "d_a")) c->n_order = 2;
        else if (!std::strcmp(r, "d_a_t")) c->n_order = 3;
        else c->fea_delta = 0;
        return CTU_OK;
    }
    if (opt == "-fea_trap") {
        // src/io/opts.cc:694-704: ignored after -fea_delta; reuses the first delta stage with window (N-1)/2
        if (r && !c->fea_delta) {
            c->fea_trap = 1;
            c->trap_win = std::atoi(r);
            c->fea_delta = 1;
            c->n_order = 1;
            c->d_win = (c->trap_win - 1) / 2;
        }
        return CTU_OK;
    }
    if (opt == "-format_in") {
        // the sample decoders (raw, alaw, mulaw, wave) belong to the host; "htk" switches the library to feature input
        if (r) c->fea_in = !std::strcmp(r, "htk");
        return CTU_OK;
    }
    if (opt == "-filters") { if (r) put(c->filters, r, CTU_FBDEF); return CTU_OK; }
    if (opt == "-nfeacoefs") { if (r) c->nfeacoefs = std::atoi(r); return CTU_OK; }
    for (const Opt &o : kOpts) {
        if (opt != o.name) continue;
        if (o.kind == K_FLAG_IGNORE || o.kind == K_IGNORE) return CTU_OK;
        if (!r) return CTU_OK;  // the reference silently keeps the old value when the value is missing
        switch (o.kind) {
            case K_INT: *reinterpret_cast<int32_t *>(base + o.off) = std::atoi(r); break;
            case K_DBL: *reinterpret_cast<double *>(base + o.off) = std::atof(r); break;
            case K_FLT: *reinterpret_cast<float *>(base + o.off) = (float)std::atof(r); break;
            case K_ONOFF:
                if (!std::strcmp(r, "on")) *reinterpret_cast<int32_t *>(base + o.off) = 1;
                else if (!std::strcmp(r, "off")) *reinterpret_cast<int32_t *>(base + o.off) = 0;
                break;
            case K_STR: put(base + o.off, r); break;
            default: break;
        }
        return CTU_OK;
    }
    std::string msg = std::string("OPTS: Syntax error in option \"") + l + (r ? std::string(" ") + r : std::string()) + "\".";
    return cfg_fail(msg);
}

int ctu_config_sizeof(void) { return (int)sizeof(ctu_config); }

int ctu_config_finalize(ctu_config *c) {
    if (!c) return CTU_ERR_CONFIG;
    if (c->fs == 0) return cfg_fail("OPTS: Please specify sampling rate!");
    c->window = (int)std::floor(.5 + c->window_ms / 1000. * (double)c->fs);
    c->wshift = (int)std::floor(.5 + c->wshift_ms / 1000. * (double)c->fs);
    // smallest power of two >= window, as the reference's loop computes it (src/io/opts.cc:274-278)
    c->wfft = 0;
    for (int i = 1048576; i > 4; i /= 2)
        if ((c->window / i) == 1) c->wfft = i * (1 + ((c->window % i) != 0));
    c->wfftby2 = c->wfft / 2 + 1;
    bool sig = !std::strcmp(c->format_out, "raw") || !std::strcmp(c->format_out, "wave");
    c->phase_needed = sig ? 1 : 0;
    if (!std::strcmp(c->vadmode, "burg")) c->phase_needed = 1;
    if (c->preem >= 1.0f || c->preem < 0.0f) return cfg_fail("OPTS: Preemphasis not in range <0,1)!");
    if (sig && c->fb_power) c->fb_power = 0;  // src/io/opts.cc:312-316
    if (c->window <= 0 || c->wshift <= 0 || c->wfft == 0) return cfg_fail("OPTS: bad window / shift");
    // exponential CMS: the time constant becomes the (float) smoothing coefficient (src/io/opts.cc:273-274);
    // kept in its own field so that finalising twice is harmless
    c->cms_exp_coef = (c->fea_Z_exp != -1.f) ? (float)(1.0 - (2 * c->wshift_ms) / (double)c->fea_Z_exp) : -1.f;
    return CTU_OK;
}

int ctu_config_parse(ctu_config *c, int argc, const char *const *argv) {
    // 1st the -C file, then the command line; a token starting with '-' is never a value
    for (int j = 0; j + 1 < argc; j++) {
        if (std::strcmp(argv[j], "-C")) continue;
        std::ifstream cfg(argv[j + 1]);
        if (!cfg) return cfg_fail("OPTS: Cannot open config file!");
        std::string line;
        while (std::getline(cfg, line)) {
            size_t h = line.find('#');
            if (h != std::string::npos) line.resize(h);
            std::vector<std::string> tok;
            size_t i = 0;
            while (i < line.size() && tok.size() < 2) {
                while (i < line.size() && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) i++;
                size_t b = i;
                while (i < line.size() && line[i] != ' ' && line[i] != '\t' && line[i] != '\r') i++;
                if (i > b) tok.push_back(line.substr(b, i - b));
            }
            if (tok.empty()) continue;
            int st = ctu_config_set(c, tok[0].c_str(), tok.size() > 1 ? tok[1].c_str() : nullptr);
            if (st) return st;
        }
    }
    for (int j = 0; j < argc; j++) {
        if (argv[j][0] != '-') continue;
        const char *r = nullptr;
        if (j + 1 < argc && argv[j + 1][0] != '-') r = argv[j + 1];
        int st = ctu_config_set(c, argv[j], r);
        if (st) return st;
    }
    return ctu_config_finalize(c);
}
