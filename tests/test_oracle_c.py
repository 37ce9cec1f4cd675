"""CPU: the plain-C fp64 restatement (oracle/oracle_c.c) against the fixtures recorded from the reference."""
import numpy as np

from oracle import oracle_c as OC


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_c_oracle_ce_cases(golden):
    data, manifest = golden
    n = 0
    for case in [c for c in manifest['cases'] if c['kind'] == 'ce']:
        name, kw = case['name'], case['kw']
        go = data.get(name + '/grad_out')
        r = OC.ce(data[name + '/logits'], data[name + '/labels'], tuple(case['size']), data.get(name + '/pixel_weight'),
                  kw.get('class_weight'), case.get('ac', False), case['ignore'], kw.get('reduction', 'mean'),
                  kw.get('avg_non_ignore', False), case.get('avg_factor'), kw.get('loss_weight', 1.0),
                  case['ignore'] if case['ignore'] != -100 else None, go)
        assert _rel(r['loss'], data[name + '/loss']) < 2e-6, name
        assert _rel(r['grad'], data[name + '/grad']) < 2e-5, name
        assert abs(r['acc'] - float(data[name + '/acc'][0])) < 1e-3, name
        n += 1
    assert n >= 12


def test_c_oracle_dice_cases(golden):
    data, manifest = golden
    for case in [c for c in manifest['cases'] if c['kind'] == 'dice']:
        name, kw = case['name'], case['kw']
        r = OC.dice(data[name + '/logits'], data[name + '/labels'], kw.get('class_weight'), kw.get('ignore_index', 255),
                    kw.get('smooth', 1.0), kw.get('exponent', 2.0), kw.get('loss_weight', 1.0), 'mean', case.get('avg_factor'))
        assert _rel(r['loss'], data[name + '/loss']) < 2e-6, name
        assert _rel(r['grad'], data[name + '/grad']) < 5e-5, name


def test_c_oracle_resize_and_areas(golden):
    data, manifest = golden
    for case in [c for c in manifest['cases'] if c['kind'] == 'resize']:
        name = case['name']
        y = OC.resize_bilinear(data[name + '/x'], tuple(case['size']), case['ac'])
        assert _rel(y, data[name + '/y']) < 1e-6, name
    for i in range(3):
        a = OC.areas(data['iau/pred%d' % i], data['iau/gt%d' % i], 5, 255)
        np.testing.assert_array_equal(a, data['iau/areas'][i][[0, 2, 3]].astype(np.int64))
        p = OC.argmax(data['process/logits%d' % i])
        a = OC.areas(p, data['process/gt%d' % i], 5, -1)
        np.testing.assert_array_equal(a, data['process/areas'][i][[0, 2, 3]].astype(np.int64))


def test_c_oracle_lovasz_cases(golden):
    """LovaszLoss (models/losses/lovasz_loss.py:26-312): the ATen-independent restatement against the reference's fixtures."""
    data, manifest = golden
    cases = [c for c in manifest['cases'] if c['kind'] == 'lovasz']
    assert len(cases) >= 12
    for case in cases:
        name = case['name']
        r = OC.lovasz(data[name + '/logits'], data[name + '/labels'], avg_factor=case.get('avg_factor'),
                      ignore_index=case['ignore'], grad_out=data.get(name + '/grad_out'), **case['kw'])
        assert _rel(r['loss'], data[name + '/loss']) < 1e-5, name
        assert _rel(r['grad'], data[name + '/grad']) < 1e-4, name
