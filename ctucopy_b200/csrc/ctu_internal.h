// Internal declarations shared by the host (.cc) and device (.cu) halves of
// libctucopy_b200.so.  Nothing here is part of the ABI.
#ifndef CTU_INTERNAL_H
#define CTU_INTERNAL_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ctucopy_b200.h"

// ---- filter bank designed on the host in fp64 (ctu_fb_design.cc) ------------------------
struct CtuFbDesign {
    int nb = 0, bins = 0;
    bool inld = false;            // apply ^0.33 after projection
    std::vector<double> mat;      // [nb x bins]
    std::vector<int> lo, hi;      // first / last tap of each band
};
// returns "" on success, else the reference's error text
std::string ctu_design_fb(const ctu_config &c, CtuFbDesign &out);

// ---- enumerations resolved once at ctu_create ---------------------------------------------
enum CtuFeaKind { FEA_NONE = 0, FEA_SPEC, FEA_LOGSPEC, FEA_DCTC, FEA_LPA, FEA_LPC, FEA_TRAPDCT, FEA_TDIIR };
enum CtuNrMode { NR_NONE = 0, NR_EXTEN, NR_HWSS, NR_FWSS, NR_2FWSS };
enum CtuVadSrc { VADSRC_NONE = 0, VADSRC_BURG, VADSRC_FILE };
enum CtuVadCri { VCRI_ENERGY = 0, VCRI_CEPDIST_LPC, VCRI_CEPDIST_FEA };
enum CtuVadThr { VTHR_ABSOLUTE = 0, VTHR_PERC, VTHR_ADAPT, VTHR_DYN };

#endif
