// Filter-bank design in fp64 on the host.  What it must reproduce: FB::FB and its helpers
// in the reference (src/fea/fb.cc:20-66 ctor, :100-132 scales, :134-184 PLP trapezoids,
// :186-253 definition grammar, :255-303 rect joins, :306-429 one filter, :432-447 band
// limits).  The arithmetic expressions are kept in the reference's evaluation order so the
// designed weights agree to the last bit; the structure (vectors, a token scanner, one
// pass per sub-bank) is this project's own.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ctu_internal.h"

namespace {

struct SubBank { double f_start, f_stop; int bands, first, last; };

enum Scale { S_LIN, S_BARK, S_EXPOLOG, S_MEL, S_BAD };

Scale scale_of(const char *s) {
    if (!std::strcmp(s, "lin")) return S_LIN;
    if (!std::strcmp(s, "bark")) return S_BARK;
    if (!std::strcmp(s, "expolog")) return S_EXPOLOG;
    if (!std::strcmp(s, "mel")) return S_MEL;
    return S_BAD;
}

double warp(Scale sc, double hz) {
    switch (sc) {
        case S_LIN: return hz;
        case S_BARK: return 6. * std::log(hz / 600. + std::sqrt((hz / 600.) * (hz / 600.) + 1.));
        case S_EXPOLOG:
            if (hz <= 2000) return 700. * (std::pow(10., hz / 3988.) - 1.);
            return 2595. * std::log10(1. + hz / 700.);
        case S_MEL: return 2595 * std::log10(1. + hz / 700.);
        default: return 0.;
    }
}

double unwarp_mid(Scale sc, double w_mid) {
    switch (sc) {
        case S_LIN: return w_mid;
        case S_BARK: return 600 * std::sinh(w_mid / 6.);
        case S_EXPOLOG:
            if (w_mid <= 1521.4) return 3988. * std::log10(1. + (w_mid / 700.));
            return 700. * (std::pow(10., w_mid / 2595.) - 1);
        case S_MEL: return 700. * (std::pow(10., w_mid / 2595.) - 1.);
        default: return 0.;
    }
}

// Hermansky's equal-loudness curve, with the extra pole above 10 kHz sampling
double eq_loudness(double om, int fs) {
    double eqnum = om * om * om * om * (om * om + 5.68e7);
    double eqden;
    if (fs <= 10000)
        eqden = (om * om + 6.3e6) * (om * om + 6.3e6) * (om * om + 3.8e8);
    else
        eqden = (om * om + 6.3e6) * (om * om + 6.3e6) * (om * om + 3.8e8) * (om * om * om * om * om * om + 9.58e26);
    return eqnum / eqden;
}

// grammar: token = [[X-YHz:]K-L/]Nfilters, tokens separated by ','
bool scan_number(const char *&p, const char *set, double &val) {
    size_t n = std::strspn(p, set);
    val = std::atof(p);
    p += n;
    return n > 0;
}

std::string parse_definition(const char *defn, int fs, std::vector<SubBank> &banks) {
    const char *err = "FB: Filter bank specification parse error!";
    std::string s(defn);
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t comma = s.find(',', pos);
        std::string tok = s.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        pos = (comma == std::string::npos) ? s.size() + 1 : comma + 1;
        if (tok.empty()) continue;
        SubBank b{0., fs / 2., 0, 1, 0};
        const char *p = tok.c_str();
        double x = 0, y = 0, z = 0;
        scan_number(p, ".1234567890", x);
        if (*p == '-') {
            p++;
            scan_number(p, ".1234567890", y);
            if (!std::strncmp(p, "Hz:", 3)) {
                p += 3;
                b.f_start = x; b.f_stop = y;
                scan_number(p, "1234567890", z); b.first = (int)z;
                if (*p != '-') return err;
                p++;
                scan_number(p, "1234567890", z); b.last = (int)z;
                if (*p != '/') return err;
                p++;
                scan_number(p, "1234567890", z); b.bands = (int)z;
                if (std::strncmp(p, "filters", 7)) return err;
            } else if (*p == '/') {
                p++;
                b.first = (int)x; b.last = (int)y;
                scan_number(p, "1234567890", z); b.bands = (int)z;
                if (std::strncmp(p, "filters", 7)) return err;
            } else {
                return err;
            }
        } else if (!std::strncmp(p, "filters", 7)) {
            b.last = (int)x; b.bands = b.last;
        } else {
            return err;
        }
        banks.push_back(b);
    }
    return "";
}

}  // namespace

std::string ctu_design_fb(const ctu_config &c, CtuFbDesign &out) {
    const int bins = c.wfftby2;
    out.bins = bins;
    out.mat.clear(); out.lo.clear(); out.hi.clear();
    bool plp = !std::strcmp(c.fb_shape, "trapez");
    bool eqld = c.fb_eqld != 0;
    out.inld = c.fb_inld != 0;
    Scale sc = scale_of(c.fb_scale);
    if (plp) { sc = S_BARK; eqld = true; out.inld = true; }  // PLP forces Bark + eq-loudness + ^0.33
    std::vector<double> hz(bins), wz(bins);
    for (int i = 0; i < bins; i++) hz[i] = (double)i * c.fs / (double)c.wfft;
    if (sc != S_BAD) for (int i = 0; i < bins; i++) wz[i] = warp(sc, hz[i]);

    std::vector<std::vector<double>> rows;
    if (plp) {
        double maxBark = 6 * std::log(c.fs / 1200. + std::sqrt((c.fs / 1200.) * (c.fs / 1200.) + 1.));
        int nBark = (int)(std::floor(maxBark + .5));
        double step = maxBark / (double)nBark;
        for (int i = 0; i < nBark - 1; i++) {
            double Om = (i + 1) * step;
            double om = 3.1415926535898 * 1200 * std::sinh(Om / 6);
            double eq = eq_loudness(om, c.fs);
            std::vector<double> v(bins);
            for (int k = 0; k < bins; k++) {
                double d = wz[k] - Om;
                if (d >= -1.3 && d <= -.5) v[k] = std::pow(10., 2.5 * (0.5 + d));
                else if (std::fabs(d) < 0.5) v[k] = 1;
                else if (d >= 0.5 && d <= 2.5) v[k] = std::pow(10., 0.5 - d);
                else v[k] = 0;
                if (eqld) v[k] *= eq;
            }
            rows.push_back(v);
        }
    } else {
        std::vector<SubBank> banks;
        std::string e = parse_definition(c.fb_definition, c.fs, banks);
        if (!e.empty()) return e;
        bool rect = !std::strcmp(c.fb_shape, "rect");
        bool tri = !std::strcmp(c.fb_shape, "triang");
        if (rect) {
            // a sub-bank whose upper edge is nobody's lower edge keeps its last bin
            double df = c.fs / (double)c.wfft;
            std::vector<bool> joined(banks.size(), false);
            for (size_t i = 0; i < banks.size(); i++)
                for (size_t j = 0; j < banks.size(); j++)
                    if (banks[i].f_stop == banks[j].f_start) joined[i] = true;
            for (size_t i = 0; i < banks.size(); i++)
                if (!joined[i]) banks[i].f_stop += df;
        }
        for (const SubBank &sb : banks) {
            for (int b = sb.first; b <= sb.last; b++) {
                if (sc == S_BAD) return "FB: Unknown frequency scale!";
                double w_high = warp(sc, sb.f_stop), w_low = warp(sc, sb.f_start);
                double w_start, w_end;
                if (rect) {
                    w_start = w_low + (b - 1.) * (w_high - w_low) / (double)(sb.bands);
                    w_end = w_low + (b + 0.) * (w_high - w_low) / (double)(sb.bands);
                } else if (tri) {
                    w_start = w_low + (b - 1.) * (w_high - w_low) / (double)(sb.bands + 1);
                    w_end = w_low + (b + 1.) * (w_high - w_low) / (double)(sb.bands + 1);
                } else {
                    return "FB: Unknown filter shape!";
                }
                double eq = 1.;
                if (eqld) {
                    double w_mid = w_start + (w_end - w_start) / 2.;
                    eq = eq_loudness(2 * 3.141592653589793 * unwarp_mid(sc, w_mid), c.fs);
                }
                std::vector<double> v(bins);
                double area = 0;
                if (rect) {
                    for (int i = 0; i < bins; i++) {
                        if (wz[i] >= w_start && wz[i] < w_end) { v[i] = 1; area++; } else v[i] = 0;
                    }
                } else {
                    for (int i = 0; i < bins; i++) {
                        if (wz[i] < w_start || wz[i] > w_end) v[i] = 0;
                        else {
                            double w_mid = w_start + (w_end - w_start) / 2.;
                            v[i] = 1. - 2. * std::fabs(w_mid - wz[i]) / (w_end - w_start);
                            area += v[i];
                        }
                    }
                }
                if (c.fb_norm) for (int i = 0; i < bins; i++) v[i] *= eq / area;
                else for (int i = 0; i < bins; i++) v[i] *= eq;
                rows.push_back(v);
                if (rows.size() == 999) return "FB: Too many filters in FB!";
            }
        }
    }
    out.nb = (int)rows.size();
    if (out.nb == 0) return "FB: Filter bank specification parse error!";
    out.mat.resize((size_t)out.nb * bins);
    for (int b = 0; b < out.nb; b++) {
        for (int k = 0; k < bins; k++) out.mat[(size_t)b * bins + k] = rows[b][k];
        // first non-zero tap ... end of the first contiguous non-zero run
        int k = 0;
        while (k < bins && rows[b][k] == 0) k++;
        if (k == bins) return "FB: empty filter in filter bank (reference would read out of bounds)";
        int lo = k;
        while (k < bins && rows[b][k] != 0) k++;
        out.lo.push_back(lo);
        out.hi.push_back(k - 1);
    }
    return "";
}
