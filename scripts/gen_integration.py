"""Regenerates the `build.rs` and `ffi.rs` listings of INTEGRATION.md (sections 1 and 2) from the crate
source in rust/particular-cuda/, so that the document cannot drift from the files
(tests/test_abi_host.py::test_integration_md_lists_the_real_files checks it).
Usage: python scripts/gen_integration.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOC = os.path.join(ROOT, "INTEGRATION.md")
FILES = {"build.rs": os.path.join(ROOT, "rust", "particular-cuda", "build.rs"),
         "ffi.rs": os.path.join(ROOT, "rust", "particular-cuda", "src", "ffi.rs")}


def render(doc: str) -> str:
    for name, path in FILES.items():
        body = open(path).read().rstrip("\n")
        pat = re.compile(r"(<!-- BEGIN GENERATED: %s -->\n)(.*?)(<!-- END GENERATED: %s -->)" % (name, name), re.S)
        assert pat.search(doc), f"INTEGRATION.md lacks the generated block for {name}"
        doc = pat.sub(lambda m: m.group(1) + "```rust\n" + body + "\n```\n" + m.group(3), doc)
    return doc


if __name__ == "__main__":
    cur = open(DOC).read()
    new = render(cur)
    if "--check" in sys.argv:
        sys.exit(0 if cur == new else 1)
    open(DOC, "w").write(new)
