#!/usr/bin/env python3
"""Fixture generator (runs only where /root/reference exists).

Parses the binary-only lib/beagle.jar of the reference and writes
tests/golden/beagle_jar_abi.json: every `native` method of
beagle.BeagleJNIWrapper with its JNI descriptor, the BeagleFlag masks and the
BeagleErrorCode values.  The JSON is the committed golden vector that
tests/test_abi_symbols.py checks our libhmsbeagle-jni.so against.
"""
import io, json, struct, sys, zipfile

JAR = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/lib/beagle.jar"


def parse_class(data):
    f = io.BytesIO(data)
    rd = lambda fmt: struct.unpack(">" + fmt, f.read(struct.calcsize(">" + fmt)))
    magic, minor, major = rd("IHH")
    assert magic == 0xCAFEBABE
    (n,) = rd("H")
    cp = [None] * n
    i = 1
    while i < n:
        (tag,) = rd("B")
        if tag == 1:
            (ln,) = rd("H"); cp[i] = ("utf8", f.read(ln).decode("utf8", "replace"))
        elif tag in (3, 4):
            cp[i] = ("i4", rd("i")[0] if tag == 3 else rd("f")[0])
        elif tag in (5, 6):
            cp[i] = ("i8", rd("q")[0] if tag == 5 else rd("d")[0]); i += 1
        elif tag in (7, 8, 16, 19, 20):
            cp[i] = ("ref", rd("H")[0])
        elif tag in (9, 10, 11, 12, 17, 18):
            cp[i] = ("ref2", rd("HH"))
        elif tag == 15:
            cp[i] = ("mh", rd("BH"))
        else:
            raise ValueError(tag)
        i += 1
    access, this_c, super_c = rd("HHH")
    (ni,) = rd("H"); f.read(2 * ni)

    def attrs():
        out = []
        (na,) = rd("H")
        for _ in range(na):
            name_i, ln = rd("HI")
            out.append((cp[name_i][1], f.read(ln)))
        return out

    members = []
    for kind in ("field", "method"):
        (cnt,) = rd("H")
        for _ in range(cnt):
            acc, name_i, desc_i = rd("HHH")
            a = attrs()
            members.append((kind, acc, cp[name_i][1], cp[desc_i][1], a))
    return cp, members


def main():
    z = zipfile.ZipFile(JAR)
    out = {"source": "lib/beagle.jar (reference, binary only)", "classes": sorted(z.namelist())}
    cp, members = parse_class(z.read("beagle/BeagleJNIWrapper.class"))
    out["natives"] = [
        {"name": n, "descriptor": d}
        for kind, acc, n, d, _ in members
        if kind == "method" and acc & 0x0100
    ]
    cp, members = parse_class(z.read("beagle/Beagle.class"))
    out["beagle_interface"] = [
        {"name": n, "descriptor": d} for kind, acc, n, d, _ in members if kind == "method"
    ]
    consts = {}
    for kind, acc, n, d, a in members:
        if kind == "field":
            for an, av in a:
                if an == "ConstantValue":
                    consts[n] = cp[struct.unpack(">H", av)[0]][1]
    out["beagle_constants"] = consts
    json.dump(out, open(__file__.rsplit("/", 1)[0] + "/beagle_jar_abi.json", "w"), indent=1)
    print(len(out["natives"]), "natives;", len(out["beagle_interface"]), "interface methods;", consts)


if __name__ == "__main__":
    main()
