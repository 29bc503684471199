"""CPU tests (no GPU) of the evaluation helpers around the hot path: the oracle-side restatement of the k-fold split and
the methodology table against hand-checked cases, the native methodology table against that restatement, and the
result.dat row format (Experiment.cs:144-152)."""
import numpy as np

import experiment_ref as R
import oracle as O
import recommendersystems_b200 as rs
from recommendersystems_b200.experiment import like_count, result_row
from recommendersystems_b200.rwr import EdgeType


def _tiny():
    # users 0, 1; items 2..7 with ids 102..107; user 0 likes items 2..6 (5 likes), user 1 likes 2 and 7; friendship 0<->1
    node_id = np.array([10, 11, 102, 103, 104, 105, 106, 107], np.int64)
    node_type = np.array([1, 1, 2, 2, 2, 2, 2, 2], np.int32)
    src, dst, et = [], [], []
    for u, items in ((0, [2, 3, 4, 5, 6]), (1, [2, 7])):
        for t in items:
            src += [u, t]; dst += [t, u]; et += [EdgeType.LIKE, EdgeType.LIKE]
    src += [0, 1]; dst += [1, 0]; et += [EdgeType.FRIENDSHIP, EdgeType.FRIENDSHIP]
    order = np.argsort(np.asarray(src), kind="stable")
    return dict(node_id=node_id, node_type=node_type, src=np.asarray(src, np.int32)[order], dst=np.asarray(dst, np.int32)[order],
                etype=np.asarray(et, np.int32)[order], w=np.ones(len(src)))


def test_split_like_history_by_hand():
    """DataLoader.cs:122-140: 5 likes, 2 folds -> unitSize 2; fold 0 = positions [0, 2), the last fold takes the rest."""
    ids = [106, 102, 105, 103, 104]
    train, test = R.split_like_history(ids, 2, 0)
    assert test.tolist() == [102, 103] and train.tolist() == [104, 105, 106]
    train, test = R.split_like_history(ids, 2, 1)
    assert test.tolist() == [104, 105, 106] and train.tolist() == [102, 103]
    # fewer likes than folds: unitSize 0, every fold but the last is empty
    assert R.split_like_history([7, 5], 3, 0)[1].tolist() == [] and R.split_like_history([7, 5], 3, 2)[1].tolist() == [5, 7]


def test_hold_out_removes_both_directions_and_keeps_order():
    links = _tiny()
    held, test = R.hold_out(links, [0, 1], 2, 1)
    assert test[0].tolist() == [104, 105, 106] and test[1].tolist() == [107]     # user 1: 2 likes, unit 1, last fold = [1, 2)
    pairs = set(zip(held["src"].tolist(), held["dst"].tolist()))
    for t in (4, 5, 6):
        assert (0, t) not in pairs and (t, 0) not in pairs
    for t in (2, 3):
        assert (0, t) in pairs and (t, 0) in pairs
    assert (1, 7) not in pairs and (7, 1) not in pairs and (1, 2) in pairs and (2, 1) in pairs
    assert (0, 1) in pairs and (1, 0) in pairs
    assert (np.diff(held["src"]) >= 0).all()


def test_hold_out_user1_window():
    # user 1 has 2 likes (ids 102, 107), 2 folds: unit 1 -> fold 1 = [1, 2) = {107}
    _, test = R.hold_out(_tiny(), [1], 2, 1)
    assert test[1].tolist() == [107]


def test_native_methodology_table_matches_the_restatement():
    """rwr_methodology_masks (DataLoader.cs:142-219 + Experiment.cs:84-101) against oracle/experiment_ref.py, which lists
    the switch cases independently."""
    bit = {"FRIENDSHIP": 0, "FOLLOWSHIP_ON_THIRDPARTY": 1, "AUTHORSHIP": 2, "MENTIONCOUNT": 3}
    etype_of = {"FRIENDSHIP": R.FRIENDSHIP, "FOLLOWSHIP_ON_THIRDPARTY": R.FOLLOW, "AUTHORSHIP": R.AUTHORSHIP, "MENTIONCOUNT": R.MENTION}
    assert [m.name for m in rs.Methodology][8] == "ALL" and len(rs.Methodology) == 16
    for m in range(16):
        feats, undef, zero = rs.methodology_masks(m)
        assert sorted(int(f) for f in feats) == sorted(bit[f] for f in R.FEATURES[m])
        want_undef = 0
        for f, t in etype_of.items():
            if f not in R.FEATURES[m]:
                want_undef |= 1 << t
        if m in R.RETYPE_FRIENDSHIP:
            want_undef |= 1 << R.FRIENDSHIP
        assert undef == want_undef, m
        assert zero == ((1 << R.MENTION) if ("MENTIONCOUNT" in R.FEATURES[m] and "FRIENDSHIP" not in R.FEATURES[m]) else 0), m
    assert rs.methodology_masks(15)[2] == 1 << R.MENTION and rs.methodology_masks(8)[1:] == (0, 0)
    try:
        rs.methodology_masks(16)
        raise AssertionError("methodology 16 must be rejected")
    except rs.RwrError:
        pass


def test_apply_methodology_by_hand():
    links = _tiny()
    base = R.apply_methodology(links, 0)                    # BASELINE: likes only
    assert set(base["etype"].tolist()) == {R.LIKE}
    m4 = R.apply_methodology(links, 4)                      # friendship loaded, then retyped UNDEFINED
    assert (m4["etype"] == 0).sum() == 2 and (m4["etype"] == R.FRIENDSHIP).sum() == 0 and len(m4["src"]) == len(links["src"])


def test_result_row_format_and_like_count():
    assert like_count(_tiny(), 0) == 5 and like_count(_tiny(), 1) == 2
    row = result_row(12345, 8, 10, 20, 7.0, 93, 1.2345678901234567)
    assert row.split("\t") == ["12345", "8", "10", "20", "7", "93", "0.123456789012346"]       # G15, Program.cs:41 reads 7 tokens
    from recommendersystems_b200.experiment import dotnet_double_to_string as fmt
    # .NET Framework's double.ToString(): fixed notation down to 1e-5 exclusive, then `E-05`; integers without a point
    assert [fmt(v) for v in (0.0, 1.0, 0.5, 0.0001, 0.00001, 1.5e-7, 123456789012345678.0, float("nan"))] == \
        ["0", "1", "0.5", "0.0001", "1E-05", "1.5E-07", "1.23456789012346E+17", "NaN"]


def test_reference_evaluate_arithmetic():
    h, ap, atk = R.evaluate_ranking([106, 104, 105, 9], [105, 106], 2)
    assert (h, atk) == (2, 1) and abs(ap - (1 / 1 + 2 / 3) / 2) < 1e-15               # Experiment.cs:121-128, :136
    h2, ap2 = O.evaluate([106, 104, 105, 9], [105, 106])
    assert h2 == h and ap2 == ap
