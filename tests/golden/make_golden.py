#!/usr/bin/env python
"""
Generates tests/golden/*.npz by running the REFERENCE binary (oracle/_ref/ctucopy4_O0,
built by oracle/build_ref.sh from the unmodified sources under /root/reference) on small
inputs.  Run in the build container (where /root/reference exists):

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

The inputs are stored once in inputs.npz, so nothing at test time needs /root/reference.
Every case runs ONE utterance per reference process: the *ss noise-reduction modes carry
state from one list entry to the next (src/nr/nr.cc:212-222), and parity for this path is
defined per utterance.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_runner as rr  # noqa: E402
from ctucopy_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B = ["-fs", "16000", "-format_in", "raw", "-dither", "0"]
MF = ["-preset", "mfcc", "-preem", "0.97"]
CONF15 = ["-format_out", "htk", "-endian_in", "little", "-endian_out", "little", "-w", "25", "-s", "10", "-preem", "0.97",
          "-fb_scale", "mel", "-fb_shape", "triang", "-fb_power", "on", "-fb_definition", "30filters", "-nr_mode", "none",
          "-fb_eqld", "off", "-fb_inld", "off", "-fea_kind", "dctc", "-fea_ncepcoefs", "12", "-fea_c0", "on", "-fea_E", "off",
          "-fea_lifter", "22", "-fea_rawenergy", "off"]

# name -> (args, kind of output, extra)
CASES = {
    # BASELINE config 1 (egs/conf/15 verbatim, egs/conf/01 static, and the 23-filter wording)
    "mfcc30_d_a": (B + CONF15 + ["-fea_delta", "d_a", "-d_win", "2", "-a_win", "2", "-t_win", "2"], "htk", {}),
    "mfcc30_static": (B + CONF15, "htk", {}),
    "mfcc23_d_a": (B + MF + ["-fb_definition", "23filters", "-format_out", "htk", "-fea_delta", "d_a"], "htk", {}),
    "mfcc26_d_a_t_win3": (B + MF + ["-format_out", "htk", "-fea_delta", "d_a_t", "-d_win", "3", "-a_win", "2", "-t_win", "1"], "htk", {}),
    "mfcc26_be": (B + MF + ["-format_out", "htk", "-endian_out", "big"], "htk_be", {}),
    # config 3
    "plpc_ark": (B + ["-preset", "plpc", "-format_out", "ark={ARK}"], "ark", {}),
    "plpc_htk_d": (B + ["-preset", "plpc", "-format_out", "htk", "-fea_delta", "d"], "htk", {}),
    "lpa_mel": (B + MF + ["-fea_kind", "lpa", "-format_out", "htk"], "htk", {}),
    # lpa on the PLP bank (cube-root compressed bands): the well-conditioned Levinson-Durbin case
    "plpc_lpa": (B + ["-preset", "plpc", "-fea_kind", "lpa", "-format_out", "htk"], "htk", {}),
    "lpc_mel_inld": (B + MF + ["-fea_kind", "lpc", "-fb_inld", "on", "-fb_eqld", "on", "-fea_ncepcoefs", "16", "-format_out", "htk"], "htk", {}),
    # config 5
    "trapdct_51_8": (B + ["-format_out", "htk", "-fb_definition", "23filters", "-fb_eqld", "off", "-fb_inld", "off", "-preem", "0.97",
                          "-fea_kind", "trapdct,51,8"], "htk", {}),
    "trapdct_11_4": (B + ["-format_out", "htk", "-fb_definition", "15filters", "-fb_eqld", "off", "-fb_inld", "off",
                          "-fea_kind", "trapdct,11,4"], "htk", {}),
    # spectra / filter banks
    "spec_lin_rect": (B + ["-format_out", "htk", "-fb_scale", "lin", "-fb_shape", "rect", "-fb_definition", "32filters",
                           "-fb_eqld", "off", "-fb_inld", "off", "-fb_norm", "off", "-fea_kind", "spec"], "htk", {}),
    "logspec_bark_tri": (B + ["-format_out", "htk", "-fb_scale", "bark", "-fb_shape", "triang", "-fb_definition", "20filters",
                              "-fb_eqld", "on", "-fb_inld", "off", "-fea_kind", "logspec", "-preem", "0.95"], "htk", {}),
    "spec_expolog_mag": (B + ["-format_out", "htk", "-fb_scale", "expolog", "-fb_shape", "triang",
                              "-fb_definition", "0-4000Hz:1-10/10filters,4000-8000Hz:1-5/5filters", "-fb_power", "off",
                              "-fb_eqld", "off", "-fb_inld", "on", "-fea_kind", "spec", "-remove_dc", "off"], "htk", {}),
    "spec_mel_rect_join": (B + ["-format_out", "htk", "-fb_scale", "mel", "-fb_shape", "rect",
                                "-fb_definition", "0-2000Hz:1-8/8filters,2000-8000Hz:2-6/6filters",
                                "-fb_eqld", "off", "-fb_inld", "off", "-fea_kind", "spec"], "htk", {}),
    # config 2
    "exten_raw": (B + ["-preset", "exten", "-format_out", "raw"], "raw", {}),
    "exten_wave_a1": (B + ["-preset", "exten", "-nr_a", "1", "-format_out", "wave"], "wave", {}),
    "exten_raw_a15_2510": (B + ["-preset", "exten", "-nr_a", "1.5", "-w", "25", "-s", "10", "-preem", "0.9", "-format_out", "raw"], "raw", {}),
    "mfcc_exten_d_a": (B + MF + ["-nr_mode", "exten", "-format_out", "htk", "-fea_delta", "d_a"], "htk", {}),
    "mfcc_exten_afterFB": (B + MF + ["-nr_mode", "exten", "-nr_when", "afterFB", "-nr_a", "2", "-format_out", "htk"], "htk", {}),
    # config 4a / 4b and the other spectral-subtraction flavours
    "fwss_burg_pfile": (B + MF + ["-nr_mode", "fwss", "-vad", "burg", "-nr_when", "beforeFB", "-format_out", "pfile={PFILE}"], "pfile", {}),
    "hwss_burg_a2": (B + MF + ["-nr_mode", "hwss", "-nr_a", "2", "-fb_power", "off", "-vad", "burg", "-format_out", "htk"], "htk", {}),
    "2fwss_burg": (B + MF + ["-nr_mode", "2fwss", "-vad", "burg", "-nr_initsegs", "5", "-format_out", "htk"], "htk", {}),
    # hwss in the LINEAR domain (band values of the half-wave rectified spectrum, magnitude and power): a bin that the
    # subtraction takes to 0 +- rounding contributes (almost) nothing here, whereas its logarithm is -inf or finite by the
    # sign of a rounding error -- this is where hwss can be held to the same tolerance as every other mode
    "hwss_burg_spec_mag": (B + MF + ["-nr_mode", "hwss", "-vad", "burg", "-fb_power", "off", "-fea_kind", "spec", "-format_out", "htk"], "htk", {}),
    "hwss_burg_spec_pow": (B + MF + ["-nr_mode", "hwss", "-nr_b", "1.5", "-vad", "burg", "-fea_kind", "spec", "-format_out", "htk"], "htk", {}),
    "fwss_file_afterFB_pfile": (B + MF + ["-nr_mode", "fwss", "-vad", "file={VADIN}", "-nr_when", "afterFB", "-format_out", "pfile={PFILE}"],
                                "pfile", {"ext_vad": True}),
    "fwss_burg_raw": (B + ["-w", "32", "-s", "16", "-nr_mode", "fwss", "-nr_b", "1.5", "-vad", "burg", "-format_out", "raw"], "raw", {}),
    # VAD module
    "vad_energy_perc": (B + MF + ["-format_out", "htk", "-vad_out_mode", "vad"], "htk", {"vad_out": True}),
    "vad_energy_dyn_drop": (B + MF + ["-format_out", "htk", "-vad_out_mode", "vad", "-vad_thr_mode", "dyn", "-vad_apply_mode", "drop"],
                            "htk", {"vad_out": True}),
    "vad_cepdist_lpc_adapt": (B + MF + ["-format_out", "htk", "-vad_out_mode", "vad", "-vad_thr_mode", "adapt", "-vad_cri_mode", "cepdist",
                                        "-vad_cepdist_mode", "lpc", "-vad", "burg"], "htk", {"vad_out": True}),
    "vad_cepdist_fea_f5_drop": (B + MF + ["-format_out", "htk", "-vad_out_mode", "vad", "-vad_thr_mode", "adapt", "-vad_cri_mode", "cepdist",
                                          "-vad_cepdist_mode", "fea", "-vad_filter_order", "5", "-vad_apply_mode", "drop"], "htk", {"vad_out": True}),
    "vad_absolute_f1": (B + MF + ["-format_out", "htk", "-vad_out_mode", "vad", "-vad_thr_mode", "absolute", "-vad_absolute_thr", "100",
                                  "-vad_filter_order", "1"], "htk", {"vad_out": True}),
    "vad_perc_d_a_drop": (B + MF + ["-format_out", "htk", "-vad_out_mode", "vad", "-fea_delta", "d_a", "-vad_apply_mode", "drop"],
                          "htk", {"vad_out": True}),
    # -vad_out_mode debug: the side files <vadfile>_vad0, _energy | _cepdist + _c0init, _thr and the threshold's own
    # (src/vad/vad.h:39-76, src/vad/vad.cc:91-94, 212-218, 375-382, 456-467, 565-576): native doubles / '0','1' characters
    "vaddbg_perc": (B + MF + ["-format_out", "htk", "-vad_out_mode", "debug"], "htk", {"vad_out": True, "vad_debug": True}),
    "vaddbg_dyn_drop_f5": (B + MF + ["-format_out", "htk", "-vad_out_mode", "debug", "-vad_thr_mode", "dyn", "-vad_apply_mode", "drop",
                                     "-vad_filter_order", "5"], "htk", {"vad_out": True, "vad_debug": True}),
    "vaddbg_adapt_cepdist_lpc": (B + MF + ["-format_out", "htk", "-vad_out_mode", "debug", "-vad_thr_mode", "adapt", "-vad_cri_mode", "cepdist",
                                           "-vad_cepdist_mode", "lpc", "-vad", "burg"], "htk", {"vad_out": True, "vad_debug": True}),
    "vaddbg_abs_f1_d_a": (B + MF + ["-format_out", "htk", "-vad_out_mode", "debug", "-vad_thr_mode", "absolute", "-vad_absolute_thr", "100",
                                    "-vad_filter_order", "1", "-fea_delta", "d_a"], "htk", {"vad_out": True, "vad_debug": True}),
    # the optional _E column: every source BATCH::init_out can pick (src/io/batch.cc:98-118)
    "mfcc_E_d_a": (B + MF + ["-fea_E", "on", "-fea_delta", "d_a", "-format_out", "htk"], "htk", {}),
    "mfcc_rawE": (B + MF + ["-fea_E", "on", "-fea_rawenergy", "on", "-format_out", "htk"], "htk", {}),
    "mfcc_E_noc0": (B + MF + ["-fea_E", "on", "-fea_c0", "off", "-format_out", "htk"], "htk", {}),
    "plp_E": (B + ["-preset", "plpc", "-fea_E", "on", "-format_out", "htk"], "htk", {}),
    "logspec_E": (B + MF + ["-fea_kind", "logspec", "-fea_E", "on", "-format_out", "htk"], "htk", {}),
    "mfcc_exten_E": (B + MF + ["-nr_mode", "exten", "-fea_E", "on", "-format_out", "htk"], "htk", {}),
    "mfcc_exten_afterFB_E": (B + MF + ["-nr_mode", "exten", "-nr_when", "afterFB", "-fea_E", "on", "-format_out", "htk"], "htk", {}),
    # exponential cepstral mean subtraction (src/fea/post_impl.cc:203-209); the block version crashes the reference
    "mfcc_cms_exp_d_a": (B + MF + ["-fea_Z_exp", "500", "-fea_delta", "d_a", "-format_out", "htk"], "htk", {}),
    "plp_cms_exp": (B + ["-preset", "plpc", "-fea_Z_exp", "2000", "-format_out", "htk"], "htk", {}),
    # other sampling rates: 256-, 1024- and 2048-point frames (the same samples read at another rate)
    "mfcc8k_d_a": (["-fs", "8000"] + B[2:] + MF + ["-fea_delta", "d_a", "-format_out", "htk"], "htk", {}),
    "plp8k": (["-fs", "8000"] + B[2:] + ["-preset", "plpc", "-format_out", "htk"], "htk", {}),
    "mfcc44k_exten_E": (["-fs", "44100"] + B[2:] + MF + ["-nr_mode", "exten", "-fea_E", "on", "-format_out", "htk"], "htk", {}),
    # -dither: one glibc rand() per loaded sample (src/io/in.cc:452-455); every golden process starts its stream at srand(1)
    "mfcc_dither1_d_a": (["-fs", "16000", "-format_in", "raw", "-dither", "1.0"] + MF + ["-fea_delta", "d_a", "-format_out", "htk"], "htk", {}),
    "plp_dither4": (["-fs", "16000", "-format_in", "raw", "-dither", "4"] + ["-preset", "plpc", "-format_out", "htk"], "htk", {}),
    # -remove_dc1: the sample ring is de-meaned in place by every frame (src/io/in.cc:343-350)
    "mfcc_dc1": (B + MF + ["-remove_dc1", "on", "-format_out", "htk"], "htk", {}),
    "mfcc_dc1_w32s8": (B + ["-preset", "mfcc", "-preem", "0", "-remove_dc1", "on", "-remove_dc", "off", "-w", "32", "-s", "8", "-nr_mode", "exten",
                           "-format_out", "htk"], "htk", {}),
    # context stacking (-fea_trap N, src/fea/fea_delta.cc:166-176) and deltas of non-cepstral vectors: deltaFEA takes the first
    # fea_ncepcoefs+1 elements of whatever FEA produced
    "mfcc_trap5": (B + MF + ["-fea_trap", "5", "-format_out", "htk"], "htk", {}),
    "mfcc_trap3_E": (B + MF + ["-fea_trap", "3", "-fea_E", "on", "-format_out", "htk"], "htk", {}),
    "logspec_trap7": (B + MF + ["-fea_kind", "logspec", "-fea_trap", "7", "-format_out", "htk"], "htk", {}),
    "logspec_d_a": (B + MF + ["-fea_kind", "logspec", "-fea_delta", "d_a", "-format_out", "htk"], "htk", {}),
    "spec22_d": (B + MF + ["-fea_kind", "spec", "-fea_ncepcoefs", "22", "-fea_delta", "d", "-format_out", "htk"], "htk", {}),
    "plp_trap5_cms": (B + ["-preset", "plpc", "-fea_trap", "5", "-fea_Z_exp", "500", "-format_out", "htk"], "htk", {}),
    "mfcc_trap5_vad_drop": (B + MF + ["-fea_trap", "5", "-format_out", "htk", "-vad_out_mode", "vad", "-vad_apply_mode", "drop"],
                            "htk", {"vad_out": True}),
    # enhanced waveforms at other sampling rates: 256-, 1024- and 2048-point synthesis (the same samples read at another rate)
    "exten_raw_8k": (["-fs", "8000"] + B[2:] + ["-preset", "exten", "-format_out", "raw"], "raw", {}),
    "exten_raw_8k_2510": (["-fs", "8000"] + B[2:] + ["-preset", "exten", "-nr_a", "2", "-w", "25", "-s", "10", "-preem", "0.97", "-format_out", "raw"], "raw", {}),
    "exten_wave_44k": (["-fs", "44100"] + B[2:] + ["-preset", "exten", "-format_out", "wave"], "wave", {}),
    "resynth_raw_22k": (["-fs", "22050"] + B[2:] + ["-nr_mode", "none", "-w", "25", "-s", "10", "-preem", "0.97", "-remove_dc", "off", "-format_out", "raw"], "raw", {}),
    # the Burg detector (spectral subtraction with its own VAD, LPC cepstral-distance criterion) at other FFT sizes
    "fwss_burg_8k": (["-fs", "8000"] + B[2:] + MF + ["-nr_mode", "fwss", "-vad", "burg", "-format_out", "htk"], "htk", {}),
    "fwss_burg_raw_8k": (["-fs", "8000"] + B[2:] + ["-w", "32", "-s", "16", "-nr_mode", "fwss", "-nr_b", "1.5", "-vad", "burg", "-format_out", "raw"], "raw", {}),
    "2fwss_burg_44k": (["-fs", "44100"] + B[2:] + MF + ["-nr_mode", "2fwss", "-vad", "burg", "-nr_initsegs", "5", "-format_out", "htk"], "htk", {}),
    "vad_cepdist_lpc_8k": (["-fs", "8000"] + B[2:] + MF + ["-format_out", "htk", "-vad_out_mode", "vad", "-vad_thr_mode", "adapt", "-vad_cri_mode", "cepdist",
                                                          "-vad_cepdist_mode", "lpc", "-vad", "burg"], "htk", {"vad_out": True}),
    # the fp64 band path (noise reduction after the filter bank, LPC on a mel bank, feature-vector VAD criterion) at other FFT sizes
    "mfcc_exten_afterFB_8k": (["-fs", "8000"] + B[2:] + MF + ["-nr_mode", "exten", "-nr_when", "afterFB", "-nr_a", "2", "-fea_E", "on", "-format_out", "htk"], "htk", {}),
    "lpc_mel_inld_44k": (["-fs", "44100"] + B[2:] + MF + ["-fea_kind", "lpc", "-fb_inld", "on", "-fb_eqld", "on", "-fea_ncepcoefs", "16", "-nr_mode", "exten", "-nr_when", "afterFB",
                                                         "-format_out", "htk"], "htk", {}),
    "vad_cepdist_fea_8k": (["-fs", "8000"] + B[2:] + MF + ["-format_out", "htk", "-vad_out_mode", "vad", "-vad_thr_mode", "adapt", "-vad_cri_mode", "cepdist",
                                                          "-vad_cepdist_mode", "fea", "-fea_delta", "d"], "htk", {"vad_out": True}),
    # odd window lengths inside 512-point frames (32 ms at 11.025 kHz = 353 samples): synthesis and the Burg detector
    "exten_raw_11k": (["-fs", "11025"] + B[2:] + ["-preset", "exten", "-format_out", "raw"], "raw", {}),
    "fwss_burg_11k": (["-fs", "11025"] + B[2:] + MF + ["-w", "32", "-s", "16", "-nr_mode", "fwss", "-vad", "burg", "-format_out", "htk"], "htk", {}),
    # -fea_kind td-iir-mfcc (src/io/in.cc:281-340): 24 time-domain IIR band filters -> windowed band energies ->
    # log -> DCT.  The coefficient file is tests/golden/tdiir_filters.asc (make_tdiir_filters.py; the reference ships none).
    # egs/conf/20 verbatim (30 / 10 ms), 25 / 10 ms (window not a multiple of the shift) and 8 kHz
    "tdiir_w30s10": (["-fs", "16000", "-format_in", "raw", "-format_out", "htk", "-endian_in", "little", "-endian_out", "little", "-w", "30", "-s", "10",
                      "-nr_mode", "none", "-fea_kind", "td-iir-mfcc", "-filters", "{FILTERS}", "-fea_ncepcoefs", "12"], "htk", {}),
    "tdiir_w25s10": (["-fs", "16000", "-format_in", "raw", "-format_out", "htk", "-w", "25", "-s", "10", "-fea_kind", "td-iir-mfcc",
                      "-filters", "{FILTERS}", "-fea_ncepcoefs", "12", "-weight_of_td_iir_mfcc_bank", "1.5"], "htk", {}),
    "tdiir_8k_w32s12": (["-fs", "8000", "-format_in", "raw", "-format_out", "htk", "-w", "32", "-s", "12", "-fea_kind", "td-iir-mfcc",
                         "-filters", "{FILTERS}", "-fea_ncepcoefs", "12"], "htk", {}),
    "logspec32k_40": (["-fs", "32000"] + B[2:] + MF + ["-fea_kind", "logspec", "-fb_definition", "40filters", "-format_out", "htk"], "htk", {}),
}


CMVN_CASES = {
    # -apply_cmvn with a statistics file that does not exist yet: statistics passes, file written, features normalised
    "cmvn_3stage_d_a": (B + MF + ["-format_out", "htk", "-fea_delta", "d_a", "-apply_cmvn", "{STAT}"], [0, 1, 4, 5], ["spkA", "spkA", "spkB", "spkB"]),
    # stacked rows / deltas of spectral vectors through the three passes (statistics order: F[1..], F[0] of the vector as written)
    "cmvn_3stage_trap3": (B + MF + ["-format_out", "htk", "-fea_trap", "3", "-apply_cmvn", "{STAT}"], [0, 4, 5], ["spkA", "spkB", "spkA"]),
    "cmvn_3stage_logspec_d": (B + MF + ["-format_out", "htk", "-fea_kind", "logspec", "-fea_delta", "d", "-apply_cmvn", "{STAT}"], [0, 4, 1], ["s1", "s2", "s2"]),
    # -stat_cmvn: statistics only
    "cmvn_stat_plp": (B + ["-preset", "plpc", "-format_out", "htk", "-stat_cmvn", "{STAT}"], [0, 4, 1, 5, 2], ["s1", "s2", "s1", "s2", "s3"]),
}


# feature-file input (-format_in htk, src/io/in.cc:623-690; BATCH::BATCH src/io/batch.cc:55-60): the input of every case is
# the committed golden "mfcc30_static" (13 columns: c1..c12, c0), one file per reference process
BH = ["-fs", "16000", "-format_in", "htk", "-fea_ncepcoefs", "12"]
FEAIN_SOURCE = "mfcc30_static"
FEAIN_CASES = {
    "feain_copy": BH + ["-fea_kind", "lpc", "-format_out", "htk"],
    "feain_lpc_d_a": BH + ["-fea_kind", "lpc", "-fea_delta", "d_a", "-format_out", "htk"],
    "feain_dctc_d_a_t_noc0": BH + ["-fea_kind", "dctc", "-fea_rawenergy", "on", "-fea_delta", "d_a_t", "-fea_c0", "off", "-format_out", "htk"],
    "feain_spec_trap5": BH + ["-fea_kind", "spec", "-fea_trap", "5", "-format_out", "htk"],
    "feain_lpc_d_cms": BH + ["-fea_kind", "lpc", "-fea_delta", "d", "-fea_Z_exp", "400", "-format_out", "htk"],
    "feain_lpa_d8": BH + ["-fea_kind", "lpa", "-fea_ncepcoefs", "8", "-fea_delta", "d", "-d_win", "3", "-format_out", "htk"],
    "feain_trap3_be": BH + ["-fea_kind", "logspec", "-fea_trap", "3", "-format_out", "htk", "-endian_out", "big"],
}


# CMVN over a list of feature files (three passes / statistics only), same layout as CMVN_CASES
FEAIN_CMVN_CASES = {
    "cmvnfea_3stage_d": (BH + ["-fea_kind", "lpc", "-fea_delta", "d", "-format_out", "htk", "-apply_cmvn", "{STAT}"], [0, 1, 4, 5], ["spkA", "spkA", "spkB", "spkB"]),
    "cmvnfea_3stage_copy": (BH + ["-fea_kind", "lpc", "-format_out", "htk", "-apply_cmvn", "{STAT}"], [0, 4, 1], ["s1", "s2", "s1"]),
    "cmvnfea_stat_trap3": (BH + ["-fea_kind", "spec", "-fea_trap", "3", "-format_out", "htk", "-stat_cmvn", "{STAT}"], [0, 4, 5], ["a", "b", "a"]),
}


# 8-bit G.711 input (-format_in alaw | mulaw): the codes are the standard inputs encoded to the nearest table value
G711_CASES = {
    "g711_alaw_mfcc_8k": ["-fs", "8000", "-format_in", "alaw", "-dither", "0"] + MF + ["-fea_delta", "d_a", "-format_out", "htk"],
    "g711_mulaw_exten_raw_8k": ["-fs", "8000", "-format_in", "mulaw", "-dither", "0", "-preset", "exten", "-format_out", "raw"],
    "g711_alaw_plp_16k": ["-fs", "16000", "-format_in", "alaw", "-dither", "0", "-preset", "plpc", "-format_out", "htk"],
}


# hwss / fwss / 2fwss over a LIST in one reference process: a file's noise estimate starts from the enhanced last frame of the
# file before it (src/nr/nr.cc:212-222, 397-408) -- the opt-in "ss_carry" mode of the library (per-utterance semantics are the
# default and what every other golden pins).  name -> (args, input indices in list order, kind)
CARRY_CASES = {
    "carry_fwss_burg": (B + MF + ["-nr_mode", "fwss", "-vad", "burg", "-format_out", "htk"], [0, 5, 4], "htk"),
    "carry_2fwss_burg_d": (B + MF + ["-nr_mode", "2fwss", "-vad", "burg", "-nr_initsegs", "5", "-fea_delta", "d", "-format_out", "htk"], [5, 0, 4, 1], "htk"),
    "carry_hwss_spec_pow": (B + MF + ["-nr_mode", "hwss", "-nr_b", "1.5", "-vad", "burg", "-fea_kind", "spec", "-format_out", "htk"], [0, 5, 4], "htk"),
    "carry_fwss_raw": (B + ["-w", "32", "-s", "16", "-nr_mode", "fwss", "-nr_b", "1.5", "-vad", "burg", "-format_out", "raw"], [0, 5, 4], "raw"),
    "carry_hwss_a2_raw_2510": (B + ["-w", "25", "-s", "10", "-preem", "0.97", "-nr_mode", "hwss", "-nr_a", "2", "-vad", "burg", "-format_out", "raw"], [4, 5, 0], "raw"),
    # the fp64 band path: FEA works in place on the band vector (dctc: its logarithm is what the next file starts from)
    "carry_fwss_file_afterFB": (B + MF + ["-nr_mode", "fwss", "-vad", "file={VADIN}", "-nr_when", "afterFB", "-format_out", "htk"], [0, 5, 4], "htk"),
    "carry_fwss_burg_8k": (["-fs", "8000"] + B[2:] + MF + ["-nr_mode", "fwss", "-vad", "burg", "-format_out", "htk"], [0, 5, 4], "htk"),
}


def g711_encode(pcm, alaw):
    """nearest code of the expansion table (ties: the smaller code); only used to make realistic test inputs"""
    import ctu_oracle as co
    table = co.g711_expand(np.arange(256, dtype=np.uint8), alaw).astype(np.int64)
    order = np.argsort(table, kind="stable")
    tv = table[order]
    x = np.asarray(pcm, dtype=np.int64)
    j = np.clip(np.searchsorted(tv, x), 1, 255)
    pick = np.where(np.abs(tv[j - 1] - x) <= np.abs(tv[j] - x), j - 1, j)
    return order[pick].astype(np.uint8)


def inputs():
    utts = [synthetic.utterance(k, 1.0) for k in (0, 1, 5, 13)]           # tone/chirp, noisy + clean
    utts.append(synthetic.utterance(2, 2.5))
    ref_sig = "/root/reference/egs/sig/SA000CB1.CS0"
    if os.path.exists(ref_sig):
        sp = np.fromfile(ref_sig, dtype="<i2")
        utts.append(sp[24000:24000 + 19200].copy())                        # 1.2 s of real speech
    utts.append(synthetic.utterance(7, 0.05)[: 400 + 160 * 2 + 37])        # 3 frames, ragged tail
    return utts


def main():
    utts = inputs()
    if not sys.argv[1:] or not os.path.exists(os.path.join(OUT, "inputs.npz")):
        np.savez_compressed(os.path.join(OUT, "inputs.npz"), **{"in%d" % i: u for i, u in enumerate(utts)})
    else:   # adding cases: the committed inputs must be the ones generated here
        z = np.load(os.path.join(OUT, "inputs.npz"))
        assert all(np.array_equal(z["in%d" % i], u) for i, u in enumerate(utts)), "inputs.npz differs from inputs()"

    rng = np.random.default_rng(7)
    names = [n for n in sys.argv[1:] if n in CASES] if sys.argv[1:] else list(CASES)
    for name in names:
        args, kind, extra = CASES[name]
        outs, vads, flags_all, dbgs = [], [], [], []
        for u in utts:
            ev = None
            if extra.get("ext_vad"):
                w = 400; s = 160
                T = (len(u) - (w - s)) // s
                ev = (rng.random(T) < 0.5).astype(np.uint8)
                ev[: min(12, T)] = 0
                flags_all.append(ev)
            r = rr.run_reference(args, [u], vad_out=bool(extra.get("vad_out")), ext_vad_bytes=None if ev is None else ev.tobytes())
            assert r["returncode"] == 0, (name, r["stderr"])
            if kind == "ark":
                b = r["files"]["out0.ark"]
                outs.append(np.frombuffer(b, dtype=np.uint8))
                vads.append(np.frombuffer(r["files"]["out0.scp"], dtype=np.uint8))
            elif kind == "pfile":
                outs.append(np.frombuffer(r["files"]["out0.pfile"], dtype=np.uint8))
            else:
                outs.append(np.frombuffer(r["outputs"][0], dtype=np.uint8))
            if extra.get("vad_out"):
                vads.append(np.frombuffer(r["vad"][0]["vad"], dtype=np.uint8))
            if extra.get("vad_debug"):
                dbgs.append({k[len("vad_"):]: np.frombuffer(v, dtype=np.uint8) for k, v in r["vad"][0].items() if k != "vad"})
        d = {"args": np.array(json.dumps(args)), "kind": np.array(kind)}
        for i, dd in enumerate(dbgs):
            for k, v in dd.items():
                d["dbg_%s_%d" % (k, i)] = v
        for i, u in enumerate(utts):
            d["out%d" % i] = outs[i]
            if vads:
                d["aux%d" % i] = vads[i]
            if flags_all:
                d["extvad%d" % i] = flags_all[i]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "ok", sum(len(o) for o in outs), "bytes")
    # CMVN over a LIST with speakers: statistics file + normalised features, one reference process
    for name, (args, idx, spk) in CMVN_CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        import subprocess, tempfile
        with tempfile.TemporaryDirectory() as d:
            for i in idx:
                utts[i].astype("<i2").tofile(os.path.join(d, "u%d.raw" % i))
            with open(os.path.join(d, "list.scp"), "w") as fh:
                for i, sp in zip(idx, spk):
                    fh.write("%s/u%d.raw %s/u%d.htk %s\n" % (d, i, d, i, sp))
            a = [x.replace("{STAT}", os.path.join(d, "cmvn.stat")) for x in args]
            pr = subprocess.run([rr.ref_binary("O0")] + a + ["-S", os.path.join(d, "list.scp")], capture_output=True, cwd=d)
            assert pr.returncode == 0, (name, pr.stderr)
            dd = {"args": np.array(json.dumps(args)), "idx": np.array(idx), "spk": np.array(json.dumps(spk)),
                  "stat": np.frombuffer(open(os.path.join(d, "cmvn.stat"), "rb").read(), dtype=np.uint8)}
            for i in idx:
                pth = os.path.join(d, "u%d.htk" % i)
                if os.path.exists(pth):
                    dd["out%d" % i] = np.frombuffer(open(pth, "rb").read(), dtype=np.uint8)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **dd)
            print(name, "ok")
    # *ss modes over a list in ONE reference process
    for name, (args, idx, kind) in CARRY_CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        import ctu_oracle as co
        o = co.parse_args([a.replace("{VADIN}", "vadin.bin") for a in args])
        pcms = [utts[i] for i in idx]
        ev = None
        if o.vadmode == "file":
            rng2 = np.random.default_rng(11)
            ev = []
            for u in pcms:
                T = co.num_frames(len(u), o)
                e = (rng2.random(T) < 0.5).astype(np.uint8)
                e[: min(12, T)] = 0
                ev.append(e)
        r = rr.run_reference(args, pcms, out_ext="out", ext_vad_bytes=None if ev is None else np.concatenate(ev).tobytes())
        assert r["returncode"] == 0 and all(x is not None for x in r["outputs"]), (name, r["stderr"])
        d = {"args": np.array(json.dumps(args)), "kind": np.array(kind), "idx": np.array(idx)}
        for j in range(len(idx)):
            d["out%d" % j] = np.frombuffer(r["outputs"][j], dtype=np.uint8)
            if ev is not None:
                d["extvad%d" % j] = ev[j]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "ok")
    # feature-file input
    for name, args in FEAIN_CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import golden_util as gu
        src = gu.Case(FEAIN_SOURCE)
        d = {"args": np.array(json.dumps(args)), "kind": np.array("htk_be" if "big" in args else "htk"), "source": np.array(FEAIN_SOURCE)}
        for i in range(len(utts)):
            r = rr.run_reference(args, [src.payload(i).astype(np.float32)])
            assert r["returncode"] == 0 and r["outputs"][0] is not None, (name, i, r["stderr"])
            d["out%d" % i] = np.frombuffer(r["outputs"][0], dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "ok")
    for name, (args, idx, spk) in FEAIN_CMVN_CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        import struct, subprocess, tempfile
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import golden_util as gu
        src = gu.Case(FEAIN_SOURCE)
        with tempfile.TemporaryDirectory() as d:
            for i in idx:
                open(os.path.join(d, "f%d.htk" % i), "wb").write(src.raw[i])
            with open(os.path.join(d, "list.scp"), "w") as fh:
                for i, sp in zip(idx, spk):
                    fh.write("%s/f%d.htk %s/g%d.htk %s\n" % (d, i, d, i, sp))
            a = [x.replace("{STAT}", os.path.join(d, "cmvn.stat")) for x in args]
            pr = subprocess.run([rr.ref_binary("O0")] + a + ["-S", os.path.join(d, "list.scp")], capture_output=True, cwd=d)
            assert pr.returncode == 0, (name, pr.stderr)
            dd = {"args": np.array(json.dumps(args)), "idx": np.array(idx), "spk": np.array(json.dumps(spk)), "source": np.array(FEAIN_SOURCE),
                  "stat": np.frombuffer(open(os.path.join(d, "cmvn.stat"), "rb").read(), dtype=np.uint8)}
            for i in idx:
                pth = os.path.join(d, "g%d.htk" % i)
                if os.path.exists(pth):
                    dd["out%d" % i] = np.frombuffer(open(pth, "rb").read(), dtype=np.uint8)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **dd)
            print(name, "ok", sorted(k for k in dd if k.startswith("out")))
    for name, args in G711_CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        alaw = "alaw" in args
        d = {"args": np.array(json.dumps(args)), "kind": np.array("raw" if args[-1] == "raw" else "htk"), "law": np.array("alaw" if alaw else "mulaw")}
        for i, u in enumerate(utts):
            codes = g711_encode(u, alaw)
            r = rr.run_reference(args, [codes])
            assert r["returncode"] == 0 and r["outputs"][0] is not None, (name, i, r["stderr"])
            d["codes%d" % i] = codes
            d["out%d" % i] = np.frombuffer(r["outputs"][0], dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "ok")
    if sys.argv[1:]:
        return       # only the named cases were asked for
    # filter-bank design goldens via the undocumented -fb_printself (src/fea/fb.cc:449-456)
    fbd = {}
    for nm, a in {
        "mel30": ["-fb_definition", "30filters", "-fb_eqld", "off"],
        "mel26_eqld": ["-fb_definition", "1-26/26filters", "-fb_eqld", "on"],
        "bark_tri20": ["-fb_scale", "bark", "-fb_definition", "20filters"],
        "plp": ["-fb_shape", "trapez"],
        "plp8k": ["-fb_shape", "trapez", "-fs", "8000"],
        "lin_rect": ["-fb_scale", "lin", "-fb_shape", "rect", "-fb_definition", "32filters", "-fb_norm", "off", "-fb_eqld", "off"],
        "expolog2": ["-fb_scale", "expolog", "-fb_definition", "0-4000Hz:1-10/10filters,4000-8000Hz:1-5/5filters", "-fb_eqld", "off"],
        "mel_rect_join": ["-fb_scale", "mel", "-fb_shape", "rect", "-fb_definition", "0-2000Hz:1-8/8filters,2000-8000Hz:2-6/6filters"],
    }.items():
        args = ["-fs", "16000", "-format_in", "raw", "-format_out", "htk", "-fea_kind", "spec", "-fb_printself"] + a
        r = rr.run_reference(args, [utts[0]])
        rows = [ln for ln in r["stderr"].splitlines() if "\t" in ln]
        m = np.array([[float(x) for x in ln.split("\t") if x.strip() != ""] for ln in rows])
        fbd[nm] = m
        fbd[nm + "_args"] = np.array(json.dumps(args))
        print("fb", nm, m.shape)
    np.savez_compressed(os.path.join(OUT, "fb_design.npz"), **fbd)


if __name__ == "__main__":
    main()
