"""The C-ABI library builds, loads, and exports every symbol include/convasr_b200.h declares.
CPU only: no compute calls."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
	src = open(os.path.join(ROOT, 'include', 'convasr_b200.h')).read()
	src = re.sub(r'/\*.*?\*/', '', src, flags = re.S)
	return sorted(set(re.findall(r'\b(cab_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_expected_surface():
	syms = declared_symbols()
	for name in ['cab_frontend_logmel', 'cab_instnorm_pack', 'cab_conv1d_fused', 'cab_grouped_conv1d', 'cab_grouped_conv1d_wgrad', 'cab_bn_multi_act_mask_fwd', 'cab_act_mask_bwd_dz', 'cab_bn_act_mask_fwd_stats', 'cab_log_softmax_argmax', 'cab_log_softmax_bwd',
				'cab_ctc_loss_fwd', 'cab_ctc_loss_bwd', 'cab_ctc_alignment', 'cab_topk_ids', 'cab_greedy_collapse', 'cab_entropy', 'cab_last_error', 'cab_abi_version', 'cab_launch_count']:
		assert name in syms


def test_library_builds_and_exports_every_declared_symbol():
	from convasr_b200 import _lib, build
	path = build.build()
	assert os.path.exists(path)
	lib = ctypes.CDLL(path)
	for name in declared_symbols():
		assert hasattr(lib, name), f'{name} declared in the header but not exported'
	loaded = _lib.load()
	assert loaded.cab_abi_version() == _lib.ABI_VERSION == 3
	# the ctypes table binds exactly the declared compute entry points
	assert sorted(_lib.SIGNATURES) == sorted(s for s in declared_symbols() if s not in _lib.INTROSPECTION)


def test_struct_layouts_match_the_header():
	from convasr_b200 import _lib
	# cab_conv_source_t: 2 pointers + 11 int32 -> 64 bytes with tail padding; epilogue struct likewise
	assert ctypes.sizeof(_lib.ConvSource) == 64
	assert _lib.ConvSource.act.offset == 0 and _lib.ConvSource.T_in.offset == 16
	assert _lib.ConvEpilogue.bias.offset == 32 and _lib.ConvEpilogue.out_T_rows.offset == 64


def test_cpu_tensors_are_rejected_loudly():
	import pytest
	import torch
	from convasr_b200 import models, ops
	with pytest.raises(RuntimeError, match = 'CUDA'):
		ops.log_softmax_argmax(torch.zeros(1, 4, 8))
	fe = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window')
	with pytest.raises(RuntimeError, match = 'no CPU fallback'):
		fe(torch.zeros(1, 800))
	m = models.Wav2Letter(64, [38], base_width = 8).eval()
	with pytest.raises(RuntimeError, match = 'no CPU fallback'):
		m(torch.zeros(1, 64, 32))


def test_ctypes_table_matches_header_prototypes():
	"""every prototype's parameter count equals the ctypes argtypes the Python side binds (the table is kept by hand);
	the structs passed by pointer have the sizes the header's layouts imply"""
	from convasr_b200 import _lib
	src = open(os.path.join(ROOT, 'include', 'convasr_b200.h')).read()
	src = re.sub(r'/\*.*?\*/', '', src, flags = re.S)
	protos = dict(re.findall(r'\bint\s+(cab_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', src, flags = re.S))
	for name, argtypes in _lib.SIGNATURES.items():
		assert name in protos, name
		params = [p for p in protos[name].split(',') if p.strip() and p.strip() != 'void']
		assert len(params) == len(argtypes), (name, len(params), len(argtypes))
	assert ctypes.sizeof(_lib.PackItem) == 72  # 5 pointers + 7 int32, padded to 8
	assert ctypes.sizeof(_lib.BnBranch) == 24
	assert _lib.ConvEpilogue.skip_frac.offset == _lib.ConvEpilogue.stats.offset + 8
	assert _lib.ConvEpilogue.skip_margin.offset == _lib.ConvEpilogue.skip_frac.offset + 12
	assert _lib.ConvEpilogue.bnr_y.offset == _lib.ConvEpilogue.skip_frac.offset + 16
	assert _lib.ConvEpilogue.bnr_C.offset == _lib.ConvEpilogue.bnr_y.offset + 32
	assert ctypes.sizeof(_lib.ConvEpilogue) == _lib.ConvEpilogue.bnr_y.offset + 48


def test_bn_backward_apply_coverage_query_is_host_only():
	"""cab_bn_bwd_apply_covers decides (on the host, no device) whether the BatchNorm-backward reduction may be folded into the
	dgrad launch: every row pitch of the model zoo is covered; a pitch the streamed kernel cannot tile keeps the separate pass"""
	from convasr_b200 import _lib
	lib = _lib.load()
	for ld in (64, 256, 384, 512, 640, 768, 896, 1024, 2048):
		assert lib.cab_bn_bwd_apply_covers(80 * 751, ld, 0) == 1, ld
	assert lib.cab_bn_bwd_apply_covers(80 * 751, 1088, 0) == 0  # 136 channel vectors: no CTA size <= 512 threads is a multiple of 32
	assert lib.cab_bn_bwd_apply_covers(80 * 751, 100, 0) == 0  # not a multiple of 8 channels
	assert 'cab_bn_bwd_apply_covers' in _lib.HOST_ONLY and 'cab_bn_bwd_apply_covers' not in _lib.trace().names
