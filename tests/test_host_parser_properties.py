"""Property tests (hypothesis) of the host-side formats: the C libsvm tokenizer in libctr_b200.so (strtof/strtol) against
the pure-Python oracle (numpy's correctly rounded float32 parse) on generated number spellings, and the TFRecord /
tf.Example writer against its reader.  CPU only (the .so's host entry points need no GPU)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import libsvm as olib
from tf_repos_b200 import input_fn
from tf_repos_b200 import tfrecord as tfr

_digits = st.text("0123456789", min_size=1, max_size=18)


@st.composite
def _number(draw):
    kind = draw(st.integers(0, 5))
    sign = draw(st.sampled_from(["", "", "-", "+"]))
    if kind == 0:
        return sign + draw(_digits)
    if kind == 1:
        return sign + draw(_digits) + "." + draw(_digits)
    if kind == 2:
        return sign + "0." + "0" * draw(st.integers(0, 30)) + draw(_digits)
    if kind == 3:
        return sign + draw(_digits)[:6] + "." + draw(_digits)[:6] + draw(st.sampled_from(["e", "E"])) + \
            draw(st.sampled_from(["", "+", "-"])) + str(draw(st.integers(0, 45)))
    if kind == 4:
        return sign + "." + draw(_digits)
    return repr(draw(st.floats(allow_nan=False, allow_infinity=False, width=32)))


@given(label=st.sampled_from(["0", "1", "0.0", "1.0", "-1"]),
       pairs=st.lists(st.tuples(st.integers(0, 2 ** 31 - 1), _number()), min_size=3, max_size=3),
       spaces=st.integers(1, 3))
@settings(max_examples=300, deadline=None)
def test_c_tokenizer_equals_python_oracle(label, pairs, spaces):
    line = label + "".join(" " * spaces + "%d:%s" % p for p in pairs)
    ids, vals, labels = input_fn._parse((line + "\n").encode(), 0, len(line) + 1, 3)
    o_ids, o_vals, o_lab = olib.decode_libsvm(line)
    assert ids.shape == (1, 3) and np.array_equal(ids[0], o_ids)
    assert np.array_equal(vals[0].view(np.uint32), o_vals.view(np.uint32)), (line, vals, o_vals)     # bit-exact floats
    assert np.float32(labels[0]) == o_lab


@given(st.dictionaries(st.text("abcdefghij_", min_size=1, max_size=12),
                       st.one_of(st.lists(st.integers(-2 ** 63, 2 ** 63 - 1), max_size=20).map(lambda v: np.asarray(v, dtype=np.int64)),
                                 st.lists(st.floats(width=32, allow_nan=False), max_size=20).map(lambda v: np.asarray(v, dtype=np.float32))),
                       max_size=8))
@settings(max_examples=150, deadline=None)
def test_example_encode_parse_roundtrip(features):
    got = tfr.parse_example(tfr.encode_example(features))
    assert set(got) == set(features)
    for k, v in features.items():
        assert got[k].dtype == v.dtype and np.array_equal(got[k], v), k
