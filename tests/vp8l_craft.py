"""Hand-made VP8L bitstreams for the corners no encoder of the reference reaches (test infrastructure).

Only "simple" prefix codes are written (one or two symbols, RFC 9649 section 3.7.2.1.1), which is enough to build
  * a picture whose meta prefix image names group numbers above 1000, most of them never used (the reference then keeps tables
    only for the groups in use, /root/reference/src/dec/vp8l_dec.c:399-424), and
  * pictures whose colour-indexing transform is not the first transform of the stream (the reference widens the rows in place,
    src/dsp/lossless.c:341-385).
The decoders under test must agree with the reference on these files; nothing here is used by the product."""
import struct


class Bits:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, value, nbits):
        assert 0 <= value < (1 << nbits) or nbits == 0
        self.acc |= value << self.n
        self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 0xff)
            self.acc >>= 8
            self.n -= 8

    def bytes(self):
        tail = bytearray(self.out)
        if self.n:
            tail.append(self.acc & 0xff)
        return bytes(tail)


class Code:
    """A simple prefix code over one or two symbols."""
    def __init__(self, *symbols):
        assert 1 <= len(symbols) <= 2
        self.symbols = list(symbols)

    def write_header(self, b):
        b.put(1, 1)                                # simple code
        b.put(len(self.symbols) - 1, 1)
        first = self.symbols[0]
        if first < 2:
            b.put(0, 1); b.put(first, 1)
        else:
            b.put(1, 1); b.put(first, 8)
        if len(self.symbols) == 2:
            b.put(self.symbols[1], 8)

    def write(self, b, symbol):
        if len(self.symbols) == 2:
            # the code gives the shorter (here: equal) lengths in symbol order: the smaller symbol is bit 0
            lo, hi = sorted(self.symbols)
            b.put(0 if symbol == lo else 1, 1)
        else:
            assert symbol == self.symbols[0]


class Group:
    """Five codes: green (+ length prefixes + cache), red, blue, alpha, distance. Literals only."""
    def __init__(self, green, red, blue, alpha):
        self.codes = [green, red, blue, alpha, Code(0)]

    def write_header(self, b):
        for c in self.codes:
            c.write_header(b)

    def write_pixel(self, b, argb):
        a, r, g, bl = (argb >> 24) & 255, (argb >> 16) & 255, (argb >> 8) & 255, argb & 255
        self.codes[0].write(b, g)
        self.codes[1].write(b, r)
        self.codes[2].write(b, bl)
        self.codes[3].write(b, a)


def group_for(pixels):
    """The group that can code exactly these ARGB values (each channel at most two distinct values)."""
    chans = [sorted({(p >> s) & 255 for p in pixels}) for s in (8, 16, 0, 24)]
    assert all(len(c) <= 2 for c in chans), chans
    return Group(*[Code(*c) for c in chans])


def write_subimage(b, pixels):
    """An image below level 0 (transform data, meta prefix image, palette): no colour cache, one group."""
    b.put(0, 1)
    g = group_for(pixels)
    g.write_header(b)
    for p in pixels:
        g.write_pixel(b, p)


def riff(payload):
    if len(payload) & 1:
        payload += b"\0"
    return b"RIFF" + struct.pack("<I", 4 + 8 + len(payload)) + b"WEBP" + b"VP8L" + struct.pack("<I", len(payload)) + payload


def replace_chunk(data, fourcc, payload):
    """`data` with the payload of its first `fourcc` chunk replaced (sizes fixed up)."""
    o = 12
    while o + 8 <= len(data):
        sz = struct.unpack("<I", data[o + 4:o + 8])[0]
        end = o + 8 + sz + (sz & 1)
        if data[o:o + 4] == fourcc:
            body = data[12:o] + fourcc + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"") + data[end:]
            return b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WEBP" + body
        o = end
    raise ValueError("no such chunk")


def picture(w, h, transforms, coded, meta=None, groups=None, alph_header=None):
    """transforms: list of ("subtract_green",) | ("palette", [argb deltas...]) | ("predictor", bits, [modes per tile]) |
    ("cross_color", bits, [codes per tile]); alph_header: see below; coded: the w' x h coded ARGB values (w' = width after bundling);
    meta: None or (precision, [group number per meta pixel]); groups: {number: Group} for every number up to the largest."""
    b = Bits()
    if alph_header is None:
        b.put(0x2f, 8); b.put(w - 1, 14); b.put(h - 1, 14); b.put(1, 1); b.put(0, 3)
    else:                                          # the payload of an ALPH chunk: one byte of method / filter / pre-processing
        b.put(alph_header, 8)                      # (alpha_dec.c:52-72), then the image stream without its own header
    xs = w
    for t in transforms:
        b.put(1, 1)
        if t[0] == "predictor" or t[0] == "cross_color":
            b.put(0 if t[0] == "predictor" else 1, 2)
            bits = t[1]
            b.put(bits - 2, 3)
            tx, ty = (xs + (1 << bits) - 1) >> bits, (h + (1 << bits) - 1) >> bits
            assert len(t[2]) == tx * ty
            write_subimage(b, [0xff000000 | (m << 8) for m in t[2]] if t[0] == "predictor" else t[2])
        elif t[0] == "subtract_green":
            b.put(2, 2)
        else:
            b.put(3, 2)
            pal = t[1]
            b.put(len(pal) - 1, 8)
            write_subimage(b, pal)
            n = len(pal)
            bits = 0 if n > 16 else 1 if n > 4 else 2 if n > 2 else 3
            xs = (xs + (1 << bits) - 1) >> bits
    b.put(0, 1)                                    # no more transforms
    b.put(0, 1)                                    # no colour cache
    assert len(coded) == xs * h, (len(coded), xs, h)
    if meta is None:
        b.put(0, 1)
        g = group_for(coded)
        g.write_header(b)
        for p in coded:
            g.write_pixel(b, p)
    else:
        precision, numbers = meta
        b.put(1, 1); b.put(precision - 2, 3)
        hx, hy = (xs + (1 << precision) - 1) >> precision, (h + (1 << precision) - 1) >> precision
        assert len(numbers) == hx * hy
        write_subimage(b, [0xff000000 | ((n >> 8) << 16) | ((n & 255) << 8) for n in numbers])
        for k in range(max(numbers) + 1):
            groups[k].write_header(b)
        for y in range(h):
            for x in range(xs):
                groups[numbers[(y >> precision) * hx + (x >> precision)]].write_pixel(b, coded[y * xs + x])
    return riff(b.bytes()) if alph_header is None else b.bytes()


def crafted_cases(alpha_host=None):
    """[(name, file)]: whole VP8L pictures; with `alpha_host` (a 17 x 16 file with an ALPH chunk) also the same corners as ALPH chunks."""
    out = []
    plain = Group(Code(0), Code(0), Code(0), Code(0))
    # 1. group numbers far above 1000 (round 1 stopped at 4096), two of them in use: the reference remaps (tables for the used ones only)
    w, h = 16, 16
    numbers = [0 if (i + i // 4) % 2 == 0 else 5000 for i in range(16)]
    def val(n, x, y):
        g = ((x + y) & 1) * 90 + (10 if n == 0 else 33)
        return 0xff000000 | (g << 8) | (0x070009 if n == 0 else ((((x * y) & 1) + 1) << 16) | 0x55)
    coded = [val(numbers[(y >> 2) * 4 + (x >> 2)], x, y) for y in range(h) for x in range(w)]
    groups = {k: plain for k in range(5001)}
    for n in (0, 5000):
        groups[n] = group_for([coded[y * w + x] for y in range(h) for x in range(w) if numbers[(y >> 2) * 4 + (x >> 2)] == n])
    out.append(("groups_5001_two_used", picture(w, h, [], coded, (2, numbers), groups)))
    # 2. more group numbers than pixels
    coded = [0xff102030, 0xff102030, 0xff10ee30, 0xff102030]
    groups = {k: plain for k in range(8)}
    groups[7] = group_for(coded)
    out.append(("groups_more_than_pixels", picture(2, 2, [], coded, (2, [7]), groups)))
    # 3. unused numbers below every threshold (kept, not remapped)
    w, h = 8, 8
    numbers = [0, 3, 3, 0]
    coded = [0xff000000 | (((x ^ y) & 1) * 200 + numbers[(y >> 2) * 2 + (x >> 2)]) << 8 | 0x400040 for y in range(h) for x in range(w)]
    groups = {k: plain for k in range(4)}
    for n in (0, 3):
        groups[n] = group_for([coded[y * w + x] for y in range(h) for x in range(w) if numbers[(y >> 2) * 2 + (x >> 2)] == n])
    out.append(("groups_unused_kept", picture(w, h, [], coded, (2, numbers), groups)))
    # 4-7. colour indexing that is not the first transform
    pal4 = [0xff102030, 0x00100030, 0x00100030, 0x00000000]
    coded = [0xff000000 | (0xE4 if (x + y) & 1 else 0x1B) << 8 for y in range(4) for x in range(4)]
    out.append(("subtract_green_then_palette4", picture(16, 4, [("subtract_green",), ("palette", pal4)], coded)))
    out.append(("palette4_then_subtract_green", picture(16, 4, [("palette", pal4), ("subtract_green",)], coded)))     # the usual order, as a control
    pal20 = [0xff000000] + [0x00030701] * 19
    coded = [0xff000000 | (3 if (x * y) & 2 else 17) << 8 for y in range(8) for x in range(12)]
    out.append(("predictor_then_palette20", picture(12, 8, [("predictor", 2, [1 if i & 1 else 11 for i in range(6)]), ("palette", pal20)], coded)))
    w, h = 13, 5
    coded = [0xff000000 | (0xA5 if (x + y) & 1 else 0x3C) << 8 for y in range(h) for x in range((w + 7) >> 3)]
    cc = [0xff000000 | (5 << 16) | ((250 if i & 1 else 2) << 8) | 3 for i in range(((w + 3) >> 2) * ((h + 3) >> 2))]
    out.append(("cross_colour_then_palette2", picture(w, h, [("cross_color", 2, cc), ("palette", [0xff804020, 0x00112233])], coded)))
    out.append(("predictor_cross_colour_palette2_subtract_green",
                picture(w, h, [("predictor", 2, [12 if i & 1 else 5 for i in range(8)]), ("cross_color", 2, cc), ("palette", [0xff804020, 0x00112233]),
                               ("subtract_green",)], coded)))
    if alpha_host is not None:   # the same corners inside an ALPH chunk (alpha = green): method 1, no filter / gradient filter
        w, h = 17, 16
        numbers = [(0 if (i & 1) else 1100) for i in range(5 * 4)]
        coded = [0xff000000 | ((((x + y) & 1) * 60 + (9 if numbers[(y >> 2) * 5 + (x >> 2)] == 0 else 120)) << 8) for y in range(h) for x in range(w)]
        groups = {k: plain for k in range(1101)}
        for n in (0, 1100):
            groups[n] = group_for([coded[y * w + x] for y in range(h) for x in range(w) if numbers[(y >> 2) * 5 + (x >> 2)] == n])
        out.append(("alph_groups_1101", replace_chunk(alpha_host, b"ALPH", picture(w, h, [], coded, (2, numbers), groups, alph_header=1))))
        coded = [0xff000000 | (0xE4 if (x + y) & 1 else 0x1B) << 8 for y in range(h) for x in range((w + 3) >> 2)]
        palg = [0xff002000, 0x00002000, 0x00002000, 0x00000000]
        out.append(("alph_subtract_green_then_palette4",
                    replace_chunk(alpha_host, b"ALPH", picture(w, h, [("subtract_green",), ("palette", palg)], coded, alph_header=1 | (3 << 2)))))
    return out
