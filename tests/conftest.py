import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CHR22_DB = os.path.join(GOLDEN, "_chr22", "chr22_cas9ngg_database")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import ff_oracle
    ff_oracle.lib()
    return ff_oracle


@pytest.fixture(scope="session")
def chr22_db_path():
    if not os.path.exists(CHR22_DB):
        pytest.skip("chr22 database not built (run __graft_entry__.build() where /root/reference exists)")
    return CHR22_DB


@pytest.fixture(scope="module")
def ff():
    import flashfry_b200.api as api
    return api


@pytest.fixture(scope="session")
def small_db(oracle, tmp_path_factory):
    """A 400 kb random genome with a repeat family, indexed into FlashFry's own format by the oracle."""
    import helpers
    d = tmp_path_factory.mktemp("smalldb")
    contigs = helpers.random_genome(101, 200_000, repeat_unit=60, n_repeats=400, n_contigs=2)
    fa = str(d / "genome.fa")
    helpers.write_fasta(fa, contigs, lower_fraction=0.2)
    dbp = str(d / "small_cas9ngg_database")
    stats = oracle.build_database(fa, dbp, "spcas9ngg")
    db = oracle.read_database(dbp)
    return dbp, db, stats
