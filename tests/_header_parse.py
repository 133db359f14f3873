"""Parses include/tic_b200.h into {function: (argument type codes, return type)} for the ABI consistency tests."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tic_b200.h")


def parse_header(path=HEADER):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    out = {}
    for m in re.finditer(r"\b(int64_t|int|const char\s*\*)\s+(tic_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        codes = ""
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    codes += "p"
                elif a.startswith("int64_t"):
                    codes += "l"
                elif a.startswith("float"):
                    codes += "f"
                elif a.startswith("int"):
                    codes += "i"
                else:
                    raise ValueError("unparsed argument %r of %s" % (a, name))
        out[name] = (codes, "l" if ret == "int64_t" else ("s" if "char" in ret else "i"))
    return out
