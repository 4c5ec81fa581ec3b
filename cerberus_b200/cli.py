"""Minimal docopt-style parser (docopt is not installed): reads `--flag=<x>  ... [default: v]`
lines from the usage string and accepts `--flag=value` / `--flag value` / bare `--flag`."""
import re
import sys


def parse_usage(doc, argv=None, version=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    opts = {}
    for line in doc.splitlines():
        m = re.match(r"\s+(--[\w-]+)(=<[^>]*>)?\s", line + " ")
        if not m or line.lstrip().startswith("run_"):
            continue
        name, takes = m.group(1), m.group(2) is not None
        d = re.search(r"\[default: ([^\]]*)\]", line)
        opts[name] = (takes, d.group(1) if d else (None if takes else False))
    args = {k: v[1] for k, v in opts.items()}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-h", "--help"):
            print(doc)
            sys.exit(0)
        if a == "--version":
            print(version or "")
            sys.exit(0)
        key, eq, val = a.partition("=")
        if key not in opts:
            sys.stderr.write("unknown option %s\n%s" % (key, doc))
            sys.exit(1)
        if opts[key][0]:
            if not eq:
                i += 1
                if i >= len(argv):
                    sys.stderr.write("option %s needs a value\n" % key)
                    sys.exit(1)
                val = argv[i]
            args[key] = val
        else:
            args[key] = True
        i += 1
    return args
