"""TEST INFRASTRUCTURE ONLY (see oracle/flow_oracle.py header) -- a pure-Python restatement of the dataset file
format of the reference (utils/createTFRecords.py:45-56 writes it through TensorFlow, input/flow_input.py:14-30 reads
it): TFRecord framing + tf.train.Example wire format.  TensorFlow itself is absent, so this follows the
published definitions: record = uint64 length | masked_crc32c(length) | payload | masked_crc32c(payload), with
masked(c) = rotr(c, 15) + 0xa282ead8; Example{features=1}, Features{feature=1: map<string,Feature>},
Feature{bytes_list=1}, BytesList{value=1}.  CRC-32C is pinned by the RFC 3720 B.4 vectors in the tests."""
import struct


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features: dict) -> bytes:
    """features: name -> bytes (each a one-element bytes_list), in the given order."""
    body = b''
    for k, v in features.items():
        feature = _ld(1, _ld(1, v))
        body += _ld(1, _ld(1, k.encode()) + _ld(2, feature))
    return _ld(1, body)


def frame(payload: bytes) -> bytes:
    hdr = struct.pack('<Q', len(payload))
    return hdr + struct.pack('<I', masked_crc(hdr)) + payload + struct.pack('<I', masked_crc(payload))


def _read_varint(buf, p):
    v = shift = 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, p
        shift += 7


def _fields(buf):
    p = 0
    while p < len(buf):
        key, p = _read_varint(buf, p)
        wt = key & 7
        if wt == 0:
            v, p = _read_varint(buf, p)
        elif wt == 1:
            v, p = buf[p:p + 8], p + 8
        elif wt == 5:
            v, p = buf[p:p + 4], p + 4
        elif wt == 2:
            n, p = _read_varint(buf, p)
            v, p = buf[p:p + n], p + n
        else:
            raise ValueError("wire type %d" % wt)
        yield key >> 3, wt, v


def decode_example(payload: bytes) -> dict:
    out = {}
    for f, wt, feats in _fields(payload):
        if f != 1 or wt != 2:
            continue
        for f2, wt2, entry in _fields(feats):
            if f2 != 1 or wt2 != 2:
                continue
            key = val = None
            for f3, wt3, v in _fields(entry):
                if f3 == 1:
                    key = bytes(v).decode()
                elif f3 == 2:
                    for f4, _, bl in _fields(v):
                        if f4 == 1:
                            for f5, _, item in _fields(bl):
                                if f5 == 1 and val is None:
                                    val = bytes(item)
            if key is not None and val is not None:
                out[key] = val
    return out


def read_records(data: bytes, check_crc=True):
    p = 0
    while p < len(data):
        (n,) = struct.unpack_from('<Q', data, p)
        if check_crc and struct.unpack_from('<I', data, p + 8)[0] != masked_crc(data[p:p + 8]):
            raise ValueError("length crc")
        payload = data[p + 12:p + 12 + n]
        if len(payload) != n:
            raise ValueError("truncated")
        if check_crc and struct.unpack_from('<I', data, p + 12 + n)[0] != masked_crc(payload):
            raise ValueError("payload crc")
        yield payload
        p += 16 + n
