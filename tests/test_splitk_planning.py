"""Host-side planning of the split-K GEMM variant (pure C++ logic, no GPU needed).  An EMPTY k-split would leave a CTA
whose accumulator never completes (its epilogue would sit in a bounded spin until the trap), so the planner's
"no empty split" guarantee is checked exhaustively here."""
import pytest

from shallowspeed_b200 import _C


@pytest.mark.parametrize("sms", [132, 148])
def test_planner_never_leaves_a_split_empty_and_keeps_two_ctas_per_sm_in_smem(sms):
    seen_split = 0
    for m in (10, 127, 128, 1000, 2048, 4096, 8192, 16384):
        for rows in (4, 8, 32, 33, 128, 256):
            for k in (10, 123, 784, 1000, 2048, 4096, 8192, 8200):
                splits, per, grid_z, stages, smem, err = _C.splitk_plan(m, rows, k, sms, 0)
                num_kb = (k + 31) // 32
                assert err == ""
                assert splits >= 1
                if splits == 1:
                    continue
                seen_split += 1
                assert grid_z == splits
                assert (splits - 1) * per < num_kb <= splits * per          # every split owns >= 1 k-block
                assert per >= 8                                              # enough work to fill the pipeline
                tiles = ((m + 127) // 128) * ((rows + (256 if rows >= 256 else (rows + 15) // 16 * 16) - 1)
                                               // (256 if rows >= 256 else (rows + 15) // 16 * 16))
                assert tiles * 2 <= sms and tiles * splits <= 2 * sms + tiles
                assert 2 <= stages <= 6 and smem <= 113 * 1024               # two CTAs per SM fit
    assert seen_split > 20


def test_forcing_an_illegal_split_count_is_rejected_not_silently_accepted():
    # 25 k-blocks into 8 splits -> ceil = 4 per split -> only 7 non-empty splits
    splits, per, grid_z, stages, smem, err = _C.splitk_plan(128, 32, 784, 148, 8)
    assert "empty split" in err
    splits, per, grid_z, stages, smem, err = _C.splitk_plan(128, 32, 784, 148, 5)      # 5 x 5 = 25: legal
    assert err == "" and grid_z == 5 and per == 5
