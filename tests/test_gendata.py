"""CPU: the sample wire / disk format (state codec of src/ggpzero/util/state.py, the attrs-JSON of
src/ggpzero/util/attrutil.py as written by src/ggpzero/distributed/server.py:300-322)."""
import base64
import gzip
import json

import numpy as np

from galvanise_zero_b200 import gendata
from galvanise_zero_b200.defs import datadesc


def test_state_codec_round_trip_and_layout():
    # src/test/test_state.py:27-110 checks exactly this round trip
    rng = np.random.RandomState(3)
    for n in (8, 130, 29753):
        bits = tuple(int(x) for x in rng.randint(0, 2, n))
        enc = gendata.encode_state(bits)
        assert isinstance(enc, str)
        raw = base64.decodebytes(enc.encode())
        assert len(raw) == (n + 7) // 8
        assert (raw[0] >> 7) & 1 == bits[0]                   # MSB first per byte (np.packbits)
        dec = gendata.decode_state(enc)
        assert dec[:n] == bits and len(dec) == 8 * ((n + 7) // 8) and not any(dec[n:])
    assert gendata.decode_state([1, 0]) == (1, 0) and gendata.decode_state((1,)) == (1,)
    from galvanise_zero_b200.interface import buf_to_tuple_reverse_bytes
    enc = gendata.encode_state([1, 0, 1, 1, 0, 0, 0, 1, 1])
    assert tuple(int(b) for b in buf_to_tuple_reverse_bytes(base64.decodebytes(enc.encode()))) == gendata.decode_state(enc)


def _samples(n=5):
    out = []
    for i in range(n):
        out.append(datadesc.Sample(state=[1, 0, 1] * 10, prev_states=[[0, 1, 1] * 10],
                                   policies=[[(3, 0.123456789), (7, 0.876543211)], [(0, 1.0)]],
                                   final_score=[1.0, 0.0], depth=i, game_length=40 + i,
                                   match_identifier='w0_%d_1' % i, has_resigned=False,
                                   resign_false_positive=False, starting_sample_depth=2,
                                   resultant_puct_score=[0.61234567, 0.38765433], resultant_puct_visits=800))
    return out


def test_generation_file_has_the_reference_layout(tmp_path):
    path = str(tmp_path / gendata.gendata_filename('reversi', 7))
    assert path.endswith('gendata_reversi_7.json.gz')
    gen = gendata.write_gendata(path, 'reversi', 'x1_7', _samples(), num_samples=4)
    assert gen.num_samples == 4
    d = json.loads(gzip.open(path, 'rt').read())
    # attrutil.asdict_plus: {'obj': {...}, 'obj__clz__': [module, class]}; lists of attrs get __clzlist__
    assert set(d) == {'obj', 'obj__clz__'}
    assert d['obj__clz__'] == ['ggpzero.defs.datadesc', 'GenerationSamples']
    assert d['obj']['samples__clzlist__'] == ['ggpzero.defs.datadesc', 'Sample']
    assert d['obj']['game'] == 'reversi' and d['obj']['with_generation'] == 'x1_7' and d['obj']['num_samples'] == 4
    s0 = d['obj']['samples'][0]
    assert set(s0) == {a.name for a in datadesc.Sample.__attrs_attrs__}
    assert isinstance(s0['state'], str) and isinstance(s0['prev_states'][0], str)      # base64 on the wire
    assert s0['policies'][0][0] == [3, 0.12346]                                        # "%.5f" floats
    back = gendata.read_gendata(path)
    assert back.game == 'reversi' and len(back.samples) == 4
    assert gendata.decode_state(back.samples[2].state)[:30] == tuple([1, 0, 1] * 10)
    assert back.samples[1].policies[0][1] == (7, 0.87654) and back.samples[3].depth == 3


def test_generation_file_written_by_the_references_own_serialiser():
    """tests/golden/refpy_gendata_reversi.json holds a generation serialised by the reference's own
    `attrutil.attr_to_json` (imported unmodified by tests/golden/make_reference_python_golden.py --gendata): this
    build reads it, writes the same JSON document back, and its samples pack into the records they came from."""
    import json
    import os
    import numpy as np
    from galvanise_zero_b200 import bases, lookup, tables
    from galvanise_zero_b200.datacache import samples_to_records
    from galvanise_zero_b200.defs import datadesc
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, 'golden', 'refpy_gendata_reversi.json')) as f:
        g = json.load(f)
    gen = gendata.generation_from_json(g['reference_json'])
    assert gen.game == 'reversi' and gen.with_generation == 'golden_gen_7' and gen.num_samples == len(gen.samples) == 5
    assert json.loads(gendata.generation_to_json(gen)) == json.loads(g['reference_json'])
    tr = bases.GdlBasesTransformer(lookup.by_name('reversi'), datadesc.generation_for('reversi'))
    recs = np.array(g['records'], dtype=np.int32)
    sw = tables.state_words(tr.num_bases)
    back = samples_to_records(gen.samples, tr, sw, recs.shape[1])
    assert np.array_equal(back[:, 24:], recs[:, 24:])                    # states, previous states, (move, probability) pairs
    assert np.array_equal(back[:, [9, 10, 11, 12, 13]], recs[:, [9, 10, 11, 12, 13]])   # lead, n_prev, n_children, final scores
