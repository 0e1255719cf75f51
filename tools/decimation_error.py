# per-band RT60 error of each decimation limit against the oracle, several rows
import sys, numpy as np
sys.path.insert(0, '.')
from audio_analysis_b200 import rt60bands as rb

from audio_analysis_b200.synth import synth_batch_numpy
from oracle import ref_port as oracle
import audio_analysis_b200._native as nat
lib = nat.load()
SR=48000
for n in (288000, 192000):
    R = 16
    x = synth_batch_numpy(R, n, SR, first_index=100)
    want = []
    for i in range(R):
        ob = oracle.band_metrics(x[i], SR, oracle.BandParams(band_mode="third", include_t20=True, include_edt=True))
        want.append(np.array([[np.nan if v is None else v for v in ob["metrics"][b["name"]]] for b in ob["bands"]]))
    want = np.array(want)
    # T30-only calls follow the limit; a call that also asks for T20 or EDT keeps D <= 4 (bands.cu: fit_decimation_limit)
    for label, with_t20, with_edt in (("t30", False, False), ("t20", True, False), ("edt", True, True)):
        s = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=with_t20, include_edt=with_edt)
        cols = 1 + int(with_t20) + int(with_edt)
        for max_d in (1, 4, 8, 12, 16, 24):
            lib.aa_bands_set_decimation(max_d)
            res = rb.analyse_rt60_bands_batch(x, SR, [str(i) for i in range(R)], s)
            tab = np.array([[[np.nan if v is None else v for v in
                              (r.band_metrics_by_name[b.name].rt60_t30_seconds, r.band_metrics_by_name[b.name].rt60_t20_seconds,
                               r.band_metrics_by_name[b.name].edt_seconds)][:cols] for b in r.band_definitions] for r in res])
            w = want[:, :, :cols]
            err = np.abs(tab - w) / w
            ok = (err <= 5e-3) | (np.abs(tab - w) <= 1e-3) | (np.isnan(tab) & np.isnan(w))
            dec = rb.decimation_classes(res[0].band_definitions, n, SR, s.transition_width_octaves, settings=s)
            print(n, label, "max_d", max_d, "classes", sorted(set(dec)), "worst rel per metric",
                  np.array2string(np.nanmax(err, axis=(0, 1)), formatter={"float_kind": lambda v: "%.2e" % v}),
                  "outside tolerance", int((~ok).sum()), "nan mismatch", int((np.isnan(tab) != np.isnan(w)).sum()))
            print("   T30 per band max:", np.array2string(np.nanmax(err[:, :, 0], axis=0), precision=1, max_line_width=400))
lib.aa_bands_set_decimation(16)
