"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the
codecs on the measure hot path.  Each test cites the reference test it
transcribes.  No GPU needed."""
import math
import struct

import numpy as np
import pytest

from oracle import oracle as O


# ---- pkg/encoding/int_list_test.go:28-59 (chosen EncodeType per input + round trip)
@pytest.mark.parametrize("values,enc", [
    ([0, 2, 1, 3, 4], O.ENC_DELTA),
    ([0, 1, 4, 6, 9], O.ENC_DELTA_OF_DELTA),
    ([0, 0, 0, 0, 0], O.ENC_CONST),
    ([0, 1, 2, 3, 4], O.ENC_DELTA_CONST),
])
def test_int64_list_encode_type(values, enc):
    body, got_enc, first = O.int64_list_encode(values)
    assert got_enc == enc
    assert first == 0
    assert O.int64_list_decode(body, enc, first, len(values)).tolist() == values


# ---- pkg/encoding/delta_test.go:31-79,111-159 (round trips on 8 vectors, both codecs)
DELTA_VECTORS = [
    [10, 20, 30, 40, 50], [-10, -5, 0, 5, 10], [1000000, 2000000, 3000000, 4000000, 5000000],
    [5, 5, 5, 5, 5], [10, 8, 6, 4, 2], [1, 2, 3, 4, 5], [-3, -2, 0, 2, 3],
    [987654321, 123456789, 987654321, 987654321, 123456789],
]


@pytest.mark.parametrize("values", DELTA_VECTORS)
def test_delta_roundtrip(values):
    body, enc, first = O.int64_list_encode(values)
    assert first == values[0]
    assert O.int64_list_decode(body, enc, first, len(values)).tolist() == values
    # force both explicit codecs through the decoder using hand-built bodies
    deltas = [values[i] - values[i - 1] for i in range(1, len(values))]
    assert O.int64_list_decode(O.varint_encode(deltas), O.ENC_DELTA, values[0], len(values)).tolist() == values
    d1 = deltas[0]
    dd = [deltas[i] - deltas[i - 1] for i in range(1, len(deltas))]
    assert O.int64_list_decode(O.varint_encode([d1] + dd), O.ENC_DELTA_OF_DELTA, values[0], len(values)).tolist() == values


def test_int64_wraparound():
    # all list arithmetic wraps mod 2^64 (int_list.go / delta.go use Go int64)
    vals = [-(1 << 63), (1 << 63) - 1, 0, -(1 << 63), 5]
    body, enc, first = O.int64_list_encode(vals)
    assert O.int64_list_decode(body, enc, first, len(vals)).tolist() == vals


# ---- pkg/encoding/int.go:75-99 varint: zig-zag LEB128; 1-byte form for |v| < 64
def test_varint_known_bytes():
    assert O.varint_encode([0]) == b"\x00"
    assert O.varint_encode([-1]) == b"\x01"
    assert O.varint_encode([1]) == b"\x02"
    assert O.varint_encode([63]) == b"\x7e"
    assert O.varint_encode([-64]) == b"\x7f"
    assert O.varint_encode([64]) == b"\x80\x01"
    assert O.varint_encode([-65]) == b"\x81\x01"
    assert len(O.varint_encode([(1 << 63) - 1])) == 10
    assert len(O.varint_encode([-(1 << 63)])) == 10
    assert O.varuint_encode(300) == b"\xac\x02"


# ---- pkg/convert/number_test.go:35-41 order-preserving Int64ToBytes goldens
@pytest.mark.parametrize("value,expected", [
    (-100, bytes([127, 255, 255, 255, 255, 255, 255, 156])),
    (-2, bytes([127, 255, 255, 255, 255, 255, 255, 254])),
    (-1, bytes([127, 255, 255, 255, 255, 255, 255, 255])),
    (0, bytes([128, 0, 0, 0, 0, 0, 0, 0])),
    (1, bytes([128, 0, 0, 0, 0, 0, 0, 1])),
    (2, bytes([128, 0, 0, 0, 0, 0, 0, 2])),
    (100, bytes([128, 0, 0, 0, 0, 0, 0, 100])),
])
def test_conv_int64_bytes(value, expected):
    assert O.conv_int64_to_bytes(value) == expected
    assert O.conv_bytes_to_int64(expected) == value


def test_conv_int64_extremes_roundtrip():
    for v in [(1 << 63) - 1, -(1 << 63) + 1, 123456789012345, -98765432109876]:
        assert O.conv_bytes_to_int64(O.conv_int64_to_bytes(v)) == v


# ---- pkg/encoding/float_test.go:28-48 decimal-float known answers
@pytest.mark.parametrize("inp,ints,exp", [
    ([1.0, 2.0, 3.0], [1, 2, 3], 0),
    ([1.23, 4.56, 7.89], [123, 456, 789], -2),
    ([1.4999, 1.5001], [14999, 15001], -4),
    ([0.1, 0.12, 0.123], [100, 120, 123], -3),
    ([1.000000000000001, 2.100000000000002, 3.1], [1000000000000001, 2100000000000002, 3100000000000000], -15),
    ([math.copysign(0.0, -1.0)], [0], 0),
    ([1.7976931348623157e308], [17976931348623157], 292),
])
def test_float_to_decimal_known(inp, ints, exp):
    got, e = O.float64_to_decimal_list(inp)
    assert got.tolist() == ints
    assert e == exp


def _diverse(n):
    # float_test.go:114-140 generateDiverseFloats
    out = []
    for i in range(n):
        v = float(i % 100 + 1)
        k = i % 8
        out.append([v / 100, v, v * 100, -v / 100, -v, -v * 100, v / 10000, 0.0][k])
    return out


# ---- pkg/encoding/float_test.go:67-117 bit-exact round trip over the 22 sets
ROUNDTRIP_SETS = [
    [3.14], [-2.718], [0.0], [1e15], [1.23, 4.56, 7.89, 0.1, 0.123456789], [0, 1, 100, -42, 999999],
    [0.1, 0.12, 0.123, 1.0, 100.0], [-1.5, -0.007, -99.99], [1.000000000000001, 2.100000000000002, 3.1],
    [5e-10, 3.14e-5, 1e-15], [1e15, 1.5e20], [1.7976931348623157e308], [5e-324], [0, 0, 0, 0, 0],
    [1.5, -1.5, 0.003, -0.003, 100, -100], [3.14, 3.14, 3.14, 3.14],
    [0.99, 1.0, 1.01, 9.99, 10.0, 10.01, 99.99, 100.0, 100.01], [0, 0.1, 1, 1.0, 10, 10.5, 100, 100.001],
    [100, 1000, 10000, 100000], [math.copysign(0.0, -1.0), 0, 1.5, -1.5],
    [0.0000000001, 0.00000000001, 0.000000000001], _diverse(1000),
]


@pytest.mark.parametrize("idx", range(len(ROUNDTRIP_SETS)))
def test_float_roundtrip_bit_exact(idx):
    inp = [float(x) for x in ROUNDTRIP_SETS[idx]]
    ints, e = O.float64_to_decimal_list(inp)
    back = O.decimal_list_to_float64(ints, e)
    for a, b in zip(inp, back.tolist()):
        assert a == b, (a, b, e)


def test_float_non_finite_not_encodable():
    # float.go:108-110 -> column falls back to a Plain page
    for bad in [float("nan"), float("inf"), float("-inf")]:
        with pytest.raises(ValueError):
            O.float64_to_decimal_list([1.0, bad])


def test_float_scale_overflow_not_encodable():
    # float_test.go:142-206: exponent alignment that overflows int64 -> error
    with pytest.raises(ValueError):
        O.float64_to_decimal_list([1e18, 1e-18])


# ---- pkg/encoding/float_test.go:244-271 computeDivisors chunking (via decode of large negative exponents)
def test_pow10_and_divisor_chunks():
    assert O.pow10(0) == 1.0
    assert O.pow10(22) == 1e22
    assert O.pow10(308) == 1e308
    assert O.pow10(309) == float("inf")
    # exp -309 -> divisors [1e308, 1e1]; SmallestNonzeroFloat64 round trip is covered above
    got = O.decimal_list_to_float64([1], -309)
    assert got[0] == (1.0 / 1e308) / 10.0


# ---- pkg/encoding/writer_test.go:27-56 bit writer golden bytes
def test_bit_writer_golden():
    w = O.BitWriter()
    w.write_byte(0xC1)
    w.write_bool(False)
    w.write_bits(0x3F, 6)
    w.write_bool(True)
    w.write_byte(0xAC)
    w.write_bits(0x01, 1)
    w.write_bits(0x1248F, 20)
    w.flush()
    w.write_byte(0x01)
    w.write_byte(0x02)
    w.write_bits(0x0F, 4)
    w.write_byte(0x80)
    w.write_byte(0x8F)
    w.flush()
    w.write_bits(0x01, 1)
    w.write_byte(0xFF)
    w.flush()
    assert w.bytes() == bytes([0xC1, 0x7F, 0xAC, 0x89, 0x24, 0x78, 0x01, 0x02, 0xF8, 0x08, 0xF0, 0xFF, 0x80])


# ---- pkg/encoding/reader_test.go:27-45 bit reader known answers
def test_bit_reader_golden():
    data = bytes([3, 255, 0xCC, 0x1A, 0xBC, 0xDE, 0x80])
    got = O.bit_reader_script(data, [-1, 8, 4, 8, 20, 0, 0])
    assert got == [3, 255, 0xC, 0xC1, 0xABCDE, 1, 0]


# ---- pkg/encoding/bytes_test.go:48-208: nil vs empty, large (zstd) blocks
def test_bytes_block_nil_vs_empty():
    items = [None, b"", b"a", None, b"hello world", b""]
    enc = O.bytes_block_encode(items)
    assert O.bytes_block_decode(enc, len(items)) == items
    # lens block: plain [0][n+1 bytes: type tag + lens]; data block: plain
    assert enc[0] == 0 and enc[1] == 1 + len(items)
    assert list(enc[2:3 + len(items)]) == [0, 0, 1, 2, 0, 12, 1]


def test_bytes_block_zstd_path():
    items = [("value-%05d" % i).encode() for i in range(500)]
    enc = O.bytes_block_encode(items)
    assert enc[0] == 1  # lens array (501 bytes) is zstd-compressed (bytes.go:291-304)
    assert O.bytes_block_decode(enc, len(items)) == items


# ---- pkg/encoding/dictionary_test.go:28-290
def test_dictionary_roundtrip_and_limit():
    items = [b"skywalking", b"banyandb", b"hello", b"world", b"hello", b"hello", None, b"", b"hello"]
    enc = O.dictionary_encode(items)
    assert enc is not None
    assert O.dictionary_decode(enc, len(items)) == items
    # <=256 distinct values allowed, 257 is not (dictionary.go:27,43-45)
    assert O.dictionary_encode([("v%d" % i).encode() for i in range(256)]) is not None
    assert O.dictionary_encode([("v%d" % i).encode() for i in range(257)]) is None


def test_dictionary_layout_single_run():
    # one value repeated: [nValues=1][bytesBlock(values)][u32 n=2][u8 width][value=0,count]
    enc = O.dictionary_encode([b"r3"] * 1000)
    assert enc[0] == 1                       # varuint nValues
    assert enc[1:4] == bytes([0, 2, 0])      # lens block plain, 2 bytes: [type8][len+1 ...]
    assert enc[4] == 3                       # len("r3")+1
    assert enc[5:9] == bytes([0, 2]) + b"r3"  # data block plain
    assert enc[9:13] == struct.pack(">I", 2)  # 2 uint32 in the RLE stream
    assert enc[13] == 10                     # width = bits.Len32(1000)
    bits = int.from_bytes(enc[14:], "big") >> (len(enc[14:]) * 8 - 20)
    assert bits == (0 << 10) | 1000


def test_bitpack_zero_values():
    assert O.bitpack_encode([]) == b"\x00\x00\x00\x00"
    # max value 0 -> width 1 (dictionary.go:207-211)
    assert O.bitpack_encode([0, 0, 0]) == struct.pack(">I", 3) + bytes([1, 0])


# ---- banyand/measure/column_test.go:59-116,155-262: column pages per value type
def _i64cells(vals):
    return [O.conv_int64_to_bytes(v) for v in vals]


def _f64cells(vals):
    return [struct.pack(">d", v) for v in vals]


def test_int64_column_page_layout():
    page = O.column_encode(O.VT_INT64, _i64cells([5, 7, 6, 9, 4]))
    assert page[0] == O.ENC_DELTA
    assert page[1:9] == O.conv_int64_to_bytes(5)
    assert page[9:] == O.varint_encode([2, -1, 3, -5])
    assert O.column_decode(O.VT_INT64, page, 5) == _i64cells([5, 7, 6, 9, 4])


def test_float64_column_page_layout():
    vals = [1.23, 4.56, 7.89, 0.5]
    page = O.column_encode(O.VT_FLOAT64, _f64cells(vals))
    assert page[0] == O.ENC_DELTA_OF_DELTA  # 123,456,789 then 50 -> not monotone... checked below
    assert struct.unpack(">h", page[1:3])[0] == -2
    assert page[3:11] == O.conv_int64_to_bytes(123)
    assert O.column_decode(O.VT_FLOAT64, page, 4) == _f64cells(vals)


def test_float64_column_page_delta_layout():
    vals = [25.17, 24.03, 26.9, 25.0, 24.99]
    page = O.column_encode(O.VT_FLOAT64, _f64cells(vals))
    assert page[0] == O.ENC_DELTA
    assert struct.unpack(">h", page[1:3])[0] == -2
    assert page[3:11] == O.conv_int64_to_bytes(2517)
    assert page[11:] == O.varint_encode([2403 - 2517, 2690 - 2403, 2500 - 2690, 2499 - 2500])
    assert O.column_decode(O.VT_FLOAT64, page, 5) == _f64cells(vals)


def test_numeric_column_null_falls_back_to_plain():
    cells = _i64cells([1, 2]) + [None] + _i64cells([4])
    page = O.column_encode(O.VT_INT64, cells)
    assert page[0] == O.ENC_PLAIN and page[1] == O.ENC_DICTIONARY  # column.go:147-153 then :222-234
    assert O.column_decode(O.VT_INT64, page, 4) == cells


def test_float_column_not_decimal_falls_back_to_plain():
    cells = _f64cells([1.5, float("nan"), 2.5])
    page = O.column_encode(O.VT_FLOAT64, cells)
    assert page[0] == O.ENC_PLAIN
    assert O.column_decode(O.VT_FLOAT64, page, 3) == cells


def test_string_column_high_cardinality_is_plain():
    cells = [("value_%d" % i).encode() for i in range(300)]
    page = O.column_encode(O.VT_STR, cells)
    assert page[0] == O.ENC_PLAIN
    assert O.column_decode(O.VT_STR, page, len(cells)) == cells
    low = [(b"a", b"b", b"c")[i % 3] for i in range(300)]
    page = O.column_encode(O.VT_STR, low)
    assert page[0] == O.ENC_DICTIONARY
    assert O.column_decode(O.VT_STR, page, len(low)) == low


def test_zstd_roundtrip():
    data = bytes(range(256)) * 40
    assert O.zstd_decompress(O.zstd_compress(data)) == data


def test_fast_and_general_shortest_decimal_agree():
    # the oracle's fast path for short decimals must equal the general strconv-style search
    rng = np.random.default_rng(1234)
    vals = np.concatenate([
        np.round(rng.normal(25, 5, 3000), 2), np.round(rng.random(3000) * 100, 4), rng.integers(1, 9999, 2000) * 1e-12,
        np.round(np.cumsum(rng.normal(0, 0.1, 2000)), 3), rng.random(500), rng.random(500) * 1e-7,
        np.array([0.1, 0.2, 0.3, 0.1 + 0.2, 1 / 3, 2.5e-10, 123456.789, 0.007, 7 * 0.001, 5e-324, 1.7976931348623157e308]),
    ])
    got_fast, got_slow = [], []
    for v in vals.tolist():
        O.force_slow_float(False)
        try:
            got_fast.append(tuple(x for x in (O.float64_to_decimal_list([v])[0].tolist(), O.float64_to_decimal_list([v])[1])))
        except ValueError:
            got_fast.append(None)
        O.force_slow_float(True)
        try:
            got_slow.append(tuple(x for x in (O.float64_to_decimal_list([v])[0].tolist(), O.float64_to_decimal_list([v])[1])))
        except ValueError:
            got_slow.append(None)
    O.force_slow_float(False)
    assert got_fast == got_slow


def test_shortest_decimal_matches_an_independent_shortest_repr():
    # float.go:128-190 relies on strconv.AppendFloat(f, 'e', -1, 64): the shortest digits that round-trip and, among several
    # candidates of that length, the one closest to the exact value.  Python's repr() implements the same contract (David Gay),
    # so mantissa/exponent must agree for doubles of every digit count -- 16-digit values are where two candidates round-trip.
    from decimal import Decimal
    rng = np.random.default_rng(77)
    vals = np.concatenate([rng.standard_normal(3000) * 10.0 ** rng.integers(-6, 9, 3000), rng.random(1000) * 1e5,
                           np.round(rng.random(1000) * 1e4, 3), np.array([70828.40467154187, -88378.67643533742, 0.1 + 0.2, 1 / 3])])
    for slow in (False, True):
        O.force_slow_float(slow)
        try:
            for v in vals.tolist():
                sign, digits, e = Decimal(repr(v)).as_tuple()
                m = int("".join(map(str, digits)))
                while m and m % 10 == 0:
                    m //= 10
                    e += 1
                ints, exp = O.float64_to_decimal_list([v])
                assert (int(ints[0]), exp) == ((-m if sign else m), e), (v, slow)
        finally:
            O.force_slow_float(False)


# ---- pkg/encoding/dictionary_test.go:73-157 + bytes_test.go:65-128 edge cases (nil vs empty vs content, duplicates)
EDGE_CASES = [
    [None], [b""], [b"a"], [None, b""], [b"", b"hello"], [None, None, None], [b"", b"", b""],
    [None, b"", b"test", None, b"value"], [b"", b"a", b"", b"a"], [None, b"b", None, b"b"],
]


@pytest.mark.parametrize("items", EDGE_CASES)
def test_dictionary_and_bytes_block_edge_cases(items):
    enc = O.dictionary_encode(items)
    assert enc is not None and O.dictionary_decode(enc, len(items)) == items
    assert O.bytes_block_decode(O.bytes_block_encode(items), len(items)) == items
    # the same cells as a string column page (column.go:222-234 -> dictionary) and back
    page = O.column_encode(O.VT_STR, items)
    assert page[0] == O.ENC_DICTIONARY and O.column_decode(O.VT_STR, page, len(items)) == items
