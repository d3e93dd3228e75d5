"""Sample wire / disk format: what leaves the self-play path.

* state codec — ``encode_state`` / ``decode_state`` (``src/ggpzero/util/state.py:7-37``): base64 of
  ``np.packbits`` (MSB first per byte), as the worker applies to every sample before sending it
  (``src/ggpzero/distributed/worker.py:162-166``);
* generation files — ``gendata_<game>_<step>.json.gz`` (``src/ggpzero/distributed/server.py:295-322``):
  gzip of the reference's attrs-JSON (``src/ggpzero/util/attrutil.py:32-160``): every attrs object becomes a
  dict, with a sibling key ``<name>__clz__`` / ``<name>__clzlist__`` naming ``[module, class]``. The class
  names written here are the REFERENCE's (``ggpzero.defs.datadesc``), so its trainer loads the files as is.
"""
from __future__ import annotations

import base64
import gzip
import json
import time

import attr
import numpy as np

from .defs import datadesc

REF_MODULE = 'ggpzero.defs.datadesc'


def encode_state(s) -> str:
    assert isinstance(s, (list, tuple, np.ndarray))
    return base64.encodebytes(np.packbits(np.asarray(s, dtype=np.uint8)).tobytes()).decode('ascii')


def decode_state(s):
    """-> tuple of 0/1; like the reference the length is rounded up to a multiple of 8."""
    if isinstance(s, tuple):
        return s
    if isinstance(s, list):
        return tuple(s)
    raw = base64.decodebytes(s.encode('ascii') if isinstance(s, str) else s)
    return tuple(int(x) for x in np.unpackbits(np.frombuffer(raw, dtype=np.uint8)))


def encode_sample_states(sample: datadesc.Sample) -> datadesc.Sample:
    """What ``Worker.cb_from_superviser`` does to each fetched sample (worker.py:162-166)."""
    sample.state = encode_state(sample.state)
    sample.prev_states = [encode_state(s) for s in sample.prev_states]
    return sample


def _round(x):
    if isinstance(x, float):
        return float('%.5f' % x)              # json.encoder.FLOAT_REPR = "%.5f" in the reference
    if isinstance(x, (list, tuple)):
        return [_round(i) for i in x]
    return x


def sample_to_dict(s: datadesc.Sample) -> dict:
    return {k: _round(v) for k, v in attr.asdict(s, recurse=False).items()}


def generation_to_json(gen: datadesc.GenerationSamples) -> str:
    obj = attr.asdict(gen, recurse=False)
    obj['samples'] = [sample_to_dict(s) for s in gen.samples]
    if gen.samples:
        obj['samples__clzlist__'] = [REF_MODULE, 'Sample']
    return json.dumps({'obj': obj, 'obj__clz__': [REF_MODULE, 'GenerationSamples']})


def generation_from_json(buf) -> datadesc.GenerationSamples:
    d = json.loads(buf)
    mod, name = d['obj__clz__']
    if name not in ('GenerationSamples', 'Generation'):
        raise ValueError('not a generation file: %s / %s' % (mod, name))
    obj = dict(d['obj'])
    obj.pop('samples__clzlist__', None)
    samples = []
    for sd in obj.pop('samples', []):
        sd = dict(sd)
        sd['policies'] = [[tuple(p) for p in pol] for pol in sd.get('policies', [])]
        samples.append(datadesc.Sample(**sd))
    return datadesc.GenerationSamples(samples=samples, **obj)


def gendata_filename(game, step):
    return 'gendata_%s_%s.json.gz' % (game, step)


def write_gendata(path, game, with_generation, samples, num_samples=None, encode_states=True):
    """Writes one generation file; samples' states are base64-packed as they are on the wire."""
    out = []
    for s in samples:
        s = attr.evolve(s)
        if encode_states and not isinstance(s.state, str):
            encode_sample_states(s)
        out.append(s)
    n = len(out) if num_samples is None else min(len(out), num_samples)
    gen = datadesc.GenerationSamples(game=game, date_created=time.strftime('%Y/%m/%d %H:%M'),
                                     with_generation=with_generation, num_samples=n, samples=out[:n])
    with gzip.open(path, 'wt') as f:
        f.write(generation_to_json(gen))
    return gen


def read_gendata(path) -> datadesc.GenerationSamples:
    with gzip.open(path, 'rt') as f:
        return generation_from_json(f.read())
