"""TEST INFRASTRUCTURE ONLY — a small, independent reader/writer of the Java Object Serialization Stream Protocol (version 5) in Python,
written from the protocol grammar (stream := magic version contents; newObject / newArray / newClassDesc / newString / prevObject / nullReference
with handles assigned from 0x7E0000), used to cross-check bbmap_b200/csrc/wire.cpp: files written by the C++ must parse here into the expected
object graph, and streams produced here must be read by the C++.  Unlike the C++ reader this parser is generic: it follows handles (TC_REFERENCE),
superclass chains and any primitive field type, so it does not share the C++'s assumptions about what a .block / .chrom file contains."""
import gzip
import struct

MAGIC, VERSION = 0xACED, 5
TC_NULL, TC_REFERENCE, TC_CLASSDESC, TC_OBJECT, TC_STRING, TC_ARRAY, TC_ENDBLOCKDATA = 0x70, 0x71, 0x72, 0x73, 0x74, 0x75, 0x78
BASE_HANDLE = 0x7E0000
PRIM = {"B": ">b", "C": ">H", "D": ">d", "F": ">f", "I": ">i", "J": ">q", "S": ">h", "Z": ">?"}


def slurp(path):
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw


class Parser:
    def __init__(self, data):
        self.d, self.i, self.handles = data, 0, []

    def take(self, n):
        if self.i + n > len(self.d):
            raise ValueError("truncated stream")
        b = self.d[self.i:self.i + n]; self.i += n
        return b

    def u(self, fmt):
        return struct.unpack(fmt, self.take(struct.calcsize(fmt)))[0]

    def utf(self):
        return self.take(self.u(">H")).decode("utf-8")

    def stream(self):
        if self.u(">H") != MAGIC or self.u(">H") != VERSION:
            raise ValueError("bad magic/version")
        obj = self.content()
        if self.i != len(self.d):
            raise ValueError("trailing bytes after the object")
        return obj

    def new_handle(self, o):
        self.handles.append(o); return o

    def class_desc(self):
        tc = self.u(">B")
        if tc == TC_NULL:
            return None
        if tc == TC_REFERENCE:
            return self.handles[self.u(">i") - BASE_HANDLE]
        if tc != TC_CLASSDESC:
            raise ValueError("expected a class descriptor, got 0x%02x" % tc)
        desc = {"name": self.utf(), "suid": self.u(">q"), "fields": []}
        self.new_handle(desc)
        desc["flags"] = self.u(">B")
        for _ in range(self.u(">H")):
            t = chr(self.u(">B")); name = self.utf()
            sig = self.content() if t in "[L" else None
            desc["fields"].append((t, name, sig))
        if self.u(">B") != TC_ENDBLOCKDATA:
            raise ValueError("class annotations are not supported")
        desc["super"] = self.class_desc()
        return desc

    def content(self):
        tc = self.u(">B")
        if tc == TC_NULL:
            return None
        if tc == TC_REFERENCE:
            return self.handles[self.u(">i") - BASE_HANDLE]
        if tc == TC_STRING:
            return self.new_handle(self.utf())
        if tc == TC_ARRAY:
            self.i -= 0
            desc = self.class_desc()
            arr = {"class": desc["name"], "suid": desc["suid"]}
            self.new_handle(arr)
            n = self.u(">i"); t = desc["name"][1]
            if t in PRIM:
                sz = struct.calcsize(PRIM[t])
                arr["values"] = list(struct.unpack(">%d%s" % (n, PRIM[t][1]), self.take(n * sz)))
            else:
                arr["values"] = [self.content() for _ in range(n)]
            return arr
        if tc == TC_OBJECT:
            desc = self.class_desc()
            obj = {"class": desc["name"], "suid": desc["suid"], "fields": {}, "order": []}
            self.new_handle(obj)
            chain = []
            d = desc
            while d is not None:
                chain.append(d); d = d["super"]
            for d in reversed(chain):                      # superclass data first
                for t, name, _sig in d["fields"]:
                    obj["fields"][name] = self.u(PRIM[t]) if t in PRIM else self.content()
                    obj["order"].append(name)
            return obj
        raise ValueError("unsupported type code 0x%02x" % tc)


def parse_file(path):
    return Parser(slurp(path)).stream()


# ---- writer (what java.io.ObjectOutputStream.writeObject emits for these two shapes) ----
def _utf(s):
    b = s.encode("utf-8"); return struct.pack(">H", len(b)) + b


def _array_desc(name, suid):
    return bytes([TC_ARRAY, TC_CLASSDESC]) + _utf(name) + struct.pack(">Q", suid) + bytes([2]) + struct.pack(">H", 0) + bytes([TC_ENDBLOCKDATA, TC_NULL])


def int_array_stream(values):
    return struct.pack(">HH", MAGIC, VERSION) + _array_desc("[I", 0x4DBA602676EAB2A5) + struct.pack(">i", len(values)) + struct.pack(">%di" % len(values), *values)


def chromosome_array_stream(chromosome, array, min_index, max_index, strand):
    out = struct.pack(">HH", MAGIC, VERSION) + bytes([TC_OBJECT, TC_CLASSDESC]) + _utf("dna.ChromosomeArray") + struct.pack(">q", 3199182397853127842) + bytes([2])
    fields = [("I", "chromosome"), ("I", "maxIndex"), ("I", "minIndex"), ("B", "strand")]       # primitives sorted by name, then the reference field
    out += struct.pack(">H", len(fields) + 1)
    for t, n in fields:
        out += t.encode() + _utf(n)
    out += b"[" + _utf("array") + bytes([TC_STRING]) + _utf("[B")
    out += bytes([TC_ENDBLOCKDATA, TC_NULL])
    out += struct.pack(">iiib", chromosome, max_index, min_index, strand)
    out += _array_desc("[B", 0xACF317F8060854E0) + struct.pack(">i", len(array)) + bytes(array)
    return out
