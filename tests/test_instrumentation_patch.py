"""integration/instrument_population.patch must apply to the reference tree.

The patch is the route to real reference event dumps (PANSIM_EVENT_DUMP=<file> on a patched
`pansim` binary writes the PSEV records pansim_step_replay consumes). No Rust toolchain exists in
this image, so what can be checked here is that the unified diff applies cleanly to a pristine copy
of /root/reference and that the hooks land on the lines the C ABI header cites. On a machine
without the reference (the GPU box) the test is skipped."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "instrument_population.patch")
REF = "/root/reference/pansim"

needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not on this machine")


def test_patch_touches_the_documented_files_only():
    text = open(PATCH).read()
    files = sorted(set(re.findall(r"^\+\+\+ b/(\S+)", text, flags=re.M)))
    assert files == ["pansim/src/event_dump.rs", "pansim/src/lib.rs", "pansim/src/main.rs", "pansim/src/population.rs"]
    # the new module in the patch is the committed source of integration/event_dump.rs
    added = [l[1:] for l in text.split("+++ b/pansim/src/event_dump.rs", 1)[1].split("\ndiff --git", 1)[0].splitlines()
             if l.startswith("+")]
    assert "\n".join(added).strip() == open(os.path.join(ROOT, "integration", "event_dump.rs")).read().strip()


@needs_ref
def test_patch_applies_to_the_reference(tmp_path):
    work = tmp_path / "ref"
    shutil.copytree(REF, work / "pansim")
    r = subprocess.run(["git", "apply", "--check", "--verbose", PATCH], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    subprocess.run(["git", "apply", PATCH], cwd=work, check=True)
    pop = (work / "pansim" / "src" / "population.rs").read_text().splitlines()
    main = (work / "pansim" / "src" / "main.rs").read_text()
    lib = (work / "pansim" / "src" / "lib.rs").read_text()
    assert "pub mod event_dump;" in lib
    assert "pansim::event_dump::init_from_env();" in main and "pansim::event_dump::finish_generation(j as u32);" in main

    def line_of(needle, nth=0):
        hits = [i + 1 for i, l in enumerate(pop) if needle in l]
        assert len(hits) > nth, needle
        return hits[nth]

    # hooks sit where include/pansim_b200.h and INTEGRATION.md say (line numbers of the patched file,
    # a handful of lines after the unpatched ones)
    orig = open(os.path.join(REF, "src", "population.rs")).read().splitlines()

    def orig_line(needle, nth=0):
        return [i + 1 for i, l in enumerate(orig) if needle in l][nth]

    assert orig_line("let sampled_indices: Vec<usize>") == 443
    assert 443 < line_of("event_dump::set_parents(&sampled_indices);") < 455
    assert 501 <= orig_line("row[mutant_site] = new_allele;") <= 509
    assert 525 <= orig_line("row[mutant_site] = *new_allele;") <= 538
    assert 741 <= orig_line("self.pop[[row_idx, col_idx]] = value;") <= 746
    a = line_of("event_dump::push_row_mutations(false, dump_row_idx, &dumped);")
    b = line_of("event_dump::push_row_mutations(true, dump_row_idx, &dumped);")
    c = line_of("event_dump::push_transfer(dump_is_core, pop_idx, row_idx, col_idx, value);")
    assert a < b < c
    assert pop[c - 2].strip() == "self.pop[[row_idx, col_idx]] = value;"       # recorded right after the store, in apply order
    # the patch only adds instrumentation: it removes exactly the two closure headers it re-writes
    removed = [l[1:].strip() for l in open(PATCH).read().splitlines() if l.startswith("-") and not l.startswith("---")]
    assert removed == [".for_each(|mut row| {", ".for_each(|mut row| {"]
