"""CPU restatement of PyHELP's result reductions (TEST INFRASTRUCTURE; only tests/ import it).

* :func:`post_process_output` follows ``HelpManager._post_process_output``
  (reference ``pyhelp/managers.py:448-506``) on a dict of per-cell results;
* :func:`calc_area_monthly_avg` follows ``pyhelp/output.py:127-152``;
* :func:`calc_cells_yearly_avg` follows ``pyhelp/output.py:171-203``.
"""
import numpy as np

KEYS = ['precip', 'runoff', 'evapo', 'perco', 'subrun1', 'subrun2', 'rechg']


def post_process_output(output: dict, context) -> dict:
    cellnames = list(output.keys())
    years = output[cellnames[0]]['years']
    Np, Ny = len(cellnames), len(years)
    data = {key: np.zeros((Np, Ny, 12)) for key in KEYS}
    data['cid'] = cellnames
    data['years'] = years
    data['idx_nan'] = []
    for i, cellname in enumerate(cellnames):
        for key in KEYS[:-1]:
            data[key][i, :, :] = output[cellname][key]
        if np.any(np.isnan(output[cellname]['rechg'])):
            data['idx_nan'].append(i)
            data['rechg'][i, :, :] = output[cellname]['rechg']
        elif context[i] == 2:
            if np.all(output[cellname]['subrun2'] == 0):
                data['subrun1'][i, :, :] += output[cellname]['rechg']
            else:
                data['subrun2'][i, :, :] += output[cellname]['rechg']
        else:
            data['rechg'][i, :, :] = output[cellname]['rechg']
    return data


def calc_area_monthly_avg(data: dict) -> dict:
    Np = len(data['cid']) - len(data['idx_nan'])
    return {k: np.nansum(data[k], axis=0) / Np for k in KEYS}


def calc_cells_yearly_avg(data: dict, year_from=-np.inf, year_to=np.inf) -> dict:
    years = np.asarray(data['years'])
    mask = (years >= year_from) & (years <= year_to)
    return {k: np.mean(np.sum(data[k][:, mask, :], axis=2), axis=1) for k in KEYS}
