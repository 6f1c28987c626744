"""CPU oracle for the StrainFLAIR query-side abundance path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in plain Python/numpy, what the reference's ``json2csv`` and
``compute_strains_abundance`` compute (reference files
``strainflair/json2csv.py`` and ``strainflair/compute_strains_abundance.py``),
in the array (CSR) form the CUDA path consumes.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
Nothing under ``strainflair_b200/`` imports it, and the product path raises if
the CUDA library is missing instead of falling back to this code.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against *outputs of the reference itself*:
``tests/golden/make_golden.py`` imports the reference by file path in the
build container, runs it on seeded synthetic GFA/JSON/pickle inputs and commits
the inputs and the reference's outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them against this package on any
machine, and ``tests/test_oracle_vs_reference.py`` re-runs the live comparison
on fresh random cases whenever ``/root/reference`` is present.

Modules
-------
graph      GFA + cluster pickle -> graph arrays   (json2csv.py:223-319)
decode     vg-mpmap JSON -> alignment CSR         (json2csv.py:551-739)
abundance  match / classify / pass 1 / pass 2     (json2csv.py:321-363, 741-1279)
report     gene table, histograms, unassigned,
           strain profile                         (json2csv.py:149-201, 399-549;
                                                   compute_strains_abundance.py:48-72)
"""
