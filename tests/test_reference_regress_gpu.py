"""The reference's own self-checking regression tests on the device (golden vectors 3).

test_regress/t/t_*.v (t_math_*, t_order_*, t_clk_*, t_scheduling_*, t_always_*, t_mem_*, t_select_*,
t_case_*, t_array_*, t_inst_*, t_param_*, t_concat_*, t_const_*, t_opt_*, t_var_*, t_trace_*,
t_func_*, t_gen_*, ...)
drive themselves from `clk`, compare CRC sums and individual results
with constants written in the test ("if (sum !== `EXPECTED_SUM) $stop;", checkh macros) and end
with `$write("*-* All Finished *-*\\n"); $finish;`.  The 221 of them that the lite front end parses, that check themselves and that a toggling `clk` alone drives
were compiled by the reference's own passes + V3EmitCuda (oracle/Makefile REFTESTS; the emitted
models are committed under tests/reftests/designs/ and built into tests/reftests/
libvlcu_reftests.so, the test sources are not).
Here every instance of a batch must reach $finish with no $stop, no error and no fatal, and must
have printed the closing line: the expected values are the reference's, not ours."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REFTESTS = ['t_math_cmp', 't_math_concat', 't_math_concat0', 't_math_concat64', 't_math_cond_clean', 't_math_cond_huge', 't_math_div', 't_math_div0', 't_math_equal', 't_math_mul', 't_math_pow2', 't_math_pow4', 't_math_pow5', 't_math_pow6', 't_math_repl', 't_math_reverse', 't_math_shift_rep', 't_math_shift_sel', 't_math_shiftrs', 't_math_signed2', 't_math_signed_calc', 't_math_signed_wire', 't_math_strwidth', 't_math_svl2', 't_math_tri', 't_math_vgen', 't_math_width', 't_array_compare', 't_array_mda', 't_array_sel_short', 't_array_unpacked_public', 't_assign_inline', 't_case_66bits', 't_case_auto1', 't_case_deep', 't_case_dupitems', 't_case_itemwidth', 't_case_nest', 't_case_onehot', 't_case_orig', 't_case_reducer', 't_case_unique_overlap', 't_case_wild', 't_case_x', 't_clk_concat', 't_clk_concat3', 't_clk_condflop', 't_clk_dpulse', 't_clk_dsp', 't_clk_gate_ext', 't_clk_gated_1', 't_clk_gen', 't_clk_powerdn', 't_concat_large', 't_concat_opt', 't_concat_sel', 't_concat_unpack', 't_const_bad', 't_const_hi', 't_const_op_red_scope', 't_inst_dff', 't_inst_implicit', 't_inst_notunsized', 't_inst_prepost', 't_inst_signed', 't_inst_signed1', 't_inst_tree', 't_inst_wideconst', 't_mem_cond', 't_mem_fifo', 't_mem_file', 't_mem_first', 't_mem_iforder', 't_mem_multidim', 't_mem_shift', 't_order_2d', 't_order_clkinst', 't_order_comboclkloop', 't_order_comboloop', 't_order_loop_bad', 't_order_multialways', 't_param_bit_sel', 't_param_concat', 't_param_module', 't_param_repl', 't_param_seg', 't_param_sel', 't_param_width_loc_bad', 't_param_x_unique', 't_scheduling_1', 't_scheduling_5', 't_select_ascending', 't_select_bad_range2', 't_select_bound1', 't_select_bound2', 't_select_index', 't_select_lhs_oob', 't_select_lhs_oob2', 't_select_little', 't_select_little_pack', 't_select_loop', 't_select_negative', 't_select_param', 't_select_plus', 't_select_plus_mul_pow2', 't_select_plusloop', 't_select_runtime_range', 't_select_set', 't_a1_first_cc', 't_always_chg_first', 't_always_combdly', 't_always_dly', 't_always_reorder', 't_always_split', 't_always_split_rst', 't_always_splitord', 't_assert_question', 't_attr_parenstar', 't_blocking', 't_cover_assert', 't_cover_const_compare', 't_cover_sva_notflat', 't_dfg_3679', 't_dfg_oob_array_access', 't_display_concat2', 't_display_format_wide_decimal', 't_display_wide', 't_dos', 't_flag_compiler', 't_flag_make_cmake', 't_flag_make_cmake_sc', 't_func_grey', 't_init_concat', 't_initial_dlyass', 't_opt_const_red', 't_opt_if_array', 't_opt_merge_cond_bug_3409', 't_opt_redor', 't_opt_table_display', 't_opt_table_fsm', 't_opt_table_same', 't_opt_table_signed', 't_opt_table_sparse', 't_rnd', 't_threads_counter', 't_threads_crazy', 't_trace_ascendingrange', 't_trace_empty', 't_trace_max', 't_trace_scope_no_inline', 't_trace_timescale', 't_tri_cond_eqcase_with_1', 't_tri_eqcase_input', 't_unoptflat_simple', 't_unoptflat_simple_2', 't_unoptflat_simple_3', 't_var_assign_landr', 't_var_init', 't_var_outoforder', 't_var_overcmp', 't_var_overwidth_bad', 't_var_overzero', 't_var_set_link', 't_var_vec_sel', 't_vpi_cb_iter', 't_wrapper_legacy', 't_assign_func', 't_display_recurse', 't_func_check', 't_func_crc', 't_func_endian', 't_func_first', 't_func_flip', 't_func_mlog2', 't_func_noinl', 't_func_numones', 't_func_outfirst', 't_func_outp', 't_func_paramed', 't_func_plog', 't_func_range', 't_func_regfirst', 't_func_sel', 't_func_sum', 't_func_twocall', 't_func_under', 't_func_under2', 't_func_wide', 't_math_clog2', 't_math_precedence', 't_math_signed', 't_mem_func', 't_opt_inline_cfuncs_args', 't_param_shift', 't_trace_decoration', 't_unopt_combo', 't_unroll_signed', 't_clk_concat4', 't_func_gen', 't_gen_alw', 't_gen_cond_bitrange_bad', 't_gen_cond_const', 't_gen_for0', 't_gen_for_shuffle', 't_hierarchy_identifier_bad', 't_inst_mism', 't_inst_port_array', 't_math_imm', 't_mem_multiwire', 't_mod_recurse', 't_opt_inline_cfuncs', 't_param_if_blk', 't_param_unreachable', 't_var_nonamebegin', 't_EXAMPLE', 't_for_count', 't_math_eq', 't_math_red', 't_math_signed7', 't_tri_assigndly_nba', 't_wire_types']


@pytest.fixture(scope="module")
def abi():
    import ctypes
    import os
    from verilator_b200 import abi
    assert abi.lib.vlcu_device_count() > 0, "no CUDA device: GPU tests need a B200"
    # the designs register themselves with the runtime when their library is loaded
    ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reftests", "libvlcu_reftests.so"),
                mode=ctypes.RTLD_GLOBAL)
    return abi


@pytest.mark.parametrize("rand_reset", [0, 2])
@pytest.mark.parametrize("name", REFTESTS)
def test_reference_regression_test_passes_on_every_instance(abi, name, rand_reset):
    n, evals = 40, 4000
    assert name in abi.designs()
    with abi.Model(name, n, seed=5, rand_reset=rand_reset, msg_cap=64) as m:
        # one stimulus row per eval, the same for all instances: clk toggles, other inputs stay 0
        stim = np.zeros((evals, m.in_words), dtype=np.uint32)
        word = 0
        for pname, (p, direction, width, eb, elems) in sorted(m.ports.items(), key=lambda kv: kv[1][0]):
            if direction == abi.DIR_OUT:
                continue
            if pname == "clk":
                stim[:, word] = (np.arange(evals) & 1) ^ 1
            word += elems if elems > 1 else (2 if eb == 8 else 1)
        m.run("trace", evals, stim=stim, flags=abi.RUNF_STIM_BROADCAST | abi.RUNF_STOP_ON_FINISH)
        st = m.status()
        assert (st["finish_count"] >= 1).all(), (name, "did not reach $finish", st["time"][:4])
        assert (st["stop_count"] == 0).all(), (name, "$stop at line", st["first_stop_line"][:4],
                                               "time", st["first_stop_time"][:4])
        assert (st["fatal_count"] == 0).all() and (st["error_count"] == 0).all(), name
        for inst in (0, n - 1):
            text = "".join(m.format_message(msg) for msg in m.messages(inst)[0])
            assert "*-* All Finished *-*" in text, (name, text[-300:])
        if rand_reset == 0:  # same stimulus and same reset values: same finish time
            assert len(set(st["first_finish_time"].tolist())) == 1
