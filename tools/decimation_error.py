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
    R=12
    x = synth_batch_numpy(R, n, SR, first_index=100)
    want=[]
    for i in range(R):
        ob = oracle.band_metrics(x[i], SR, oracle.BandParams(band_mode="third", include_t20=True, include_edt=True))
        want.append(np.array([[np.nan if v is None else v for v in ob["metrics"][b["name"]]] for b in ob["bands"]]))
    want=np.array(want)
    s = rb.Rt60BandsAnalysisSettings(band_mode="third", include_t20=True, include_edt=True)
    for max_d in (1,2,3,4,8,16):
        lib.aa_bands_set_decimation(max_d)
        res = rb.analyse_rt60_bands_batch(x, SR, [str(i) for i in range(R)], s)
        tab = np.array([[[np.nan if v is None else v for v in (r.band_metrics_by_name[b.name].rt60_t30_seconds, r.band_metrics_by_name[b.name].rt60_t20_seconds, r.band_metrics_by_name[b.name].edt_seconds)] for b in r.band_definitions] for r in res])
        err = np.abs(tab - want) / want
        errs = np.abs(tab - want)
        ok = (err <= 5e-3) | (errs <= 1e-3) | (np.isnan(tab) & np.isnan(want))
        nanmis = int((np.isnan(tab) != np.isnan(want)).sum())
        print(n, 'max_d', max_d, 'worst rel T30/T20/EDT', np.array2string(np.nanmax(err, axis=(0, 1)), precision=2),
              'outside tolerance', int((~ok).sum()), 'nan mismatch', nanmis)
        print('   EDT per band max:', np.array2string(np.nanmax(err[:, :, 2], axis=0), precision=1, max_line_width=400))
