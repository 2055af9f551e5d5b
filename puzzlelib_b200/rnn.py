"""Recurrent layers (LSTM, GRU, ReLU / tanh RNN; uni- and bidirectional) of the B200 backend.

Mirrors the reference's `Rnn` object -- `CuDnn.Rnn` (Cuda/Source/Libs/CuDnnRnn.c:63-1100) as wrapped by
`CudaBackend.createRnn / acquireRnnParams / updateRnnParams` (Cuda/Backend.py:171-350) and driven by
`Backend/Dnn.py:299-333` and `Modules/RNN.py:122-166`: same constructor arguments, same `forward / backwardData /
backwardParams` methods, same parameter names (`wi wf wc wo ri rf rc ro` + `bw* br*` for the LSTM, `wr wi wh rr ri rh` for
the GRU, `wi ri bwi bri` for the plain RNNs) exposed as views into one flat weight blob `W`; a bidirectional layer
contributes two entries (forward, backward) to the parameter list and its outputs are concatenated along the feature axis.

B200 design: the reference hands the whole sequence to cuDNN's legacy RNN API (not even buildable against cuDNN 9,
SURVEY F6).  Here a layer direction is GEMMs on the tcgen05 engine plus one fused pointwise kernel per time step:
  forward   G = X Wcat^T for ALL steps at once ((T*B x in) x (in x G*H));  per step the recurrent projection h[t-1] Rcat^T
            (accumulated into G[t] for LSTM / RNN, kept apart for the GRU whose candidate gate scales it by r) and the
            cell kernel (bias, gates, c[t], h[t]) in one pass over B x G*H
  backward  per step the cell kernel (gate gradients, dc) and  dh = dgates[t] Rcat;  then  dX = dG Wcat  for all steps
  params    dWcat = dG^T X,  dRcat = dG[1:]^T H[:-1]  (two large GEMMs),  db = column sums of dG
The blob packing is ours (SURVEY 7 "LSTM weight blob layout"): per layer direction the gate matrices stacked in cuDNN's
linear-layer order so that all gates of a step are one GEMM; all matrices first, then all biases (bw, br), like cuDNN.
Dropout between layers is not implemented.
"""
import numpy as np

from .driver import lib, check
from .gpuarray import GPUArray

_f32 = np.dtype(np.float32)

MODE_RELU, MODE_TANH, MODE_LSTM, MODE_GRU = 0, 1, 2, 3
DIR_UNI, DIR_BI = 0, 1

# gate names in cuDNN's linear-layer order (Cuda/Backend.py:221-350)
_GATES = {MODE_RELU: ("i", ), MODE_TANH: ("i", ), MODE_LSTM: ("i", "f", "c", "o"), MODE_GRU: ("r", "i", "h")}


class RnnReserve:
	"""What the backward passes need from the forward pass (the reference's opaque `reserve` buffer)."""

	def __init__(self):
		self.cells = []           # per layer direction: dict of saved tensors
		self.outs = []            # per layer: the (T, B, ndir*H) output
		self.rands = []           # per layer boundary: the random words of the inter-layer dropout (None without dropout)
		self.grads = None         # filled by backwardData, consumed by backwardParams


class Rnn:
	def __init__(self, backend, insize, hsize, dtype, layers=1, algo=0, mode=MODE_LSTM, direction=DIR_UNI, dropout=0.0, seed=0,
				 batchsize=0):
		if np.dtype(dtype) != _f32:
			raise NotImplementedError("recurrent layers are float32 only (the reference creates them with np.float32, Dnn.py:301)")
		if mode not in _GATES:
			raise ValueError("invalid rnn mode %s" % mode)
		if direction not in (DIR_UNI, DIR_BI):
			raise ValueError("invalid rnn direction %s" % direction)
		if not 0.0 <= dropout < 1.0:
			raise ValueError("invalid rnn dropout probability %s" % dropout)

		self.backend = backend
		self.insize, self.hsize, self.layers = int(insize), int(hsize), int(layers)
		self.dtype, self.algo, self.mode, self.direction = _f32, algo, mode, direction
		self.dropout, self.seed, self.batchsize = dropout, seed, batchsize
		self.rng = None
		self.ngates = len(_GATES[mode])
		self.ndir = 2 if direction == DIR_BI else 1

		# blob layout: [cell: Wcat (G*H, in_l) | Rcat (G*H, H)] ... [cell: bw (G*H) | br (G*H)] ...; cell = layer * ndir + dir
		H, G = self.hsize, self.ngates
		self.matOffsets, self.biasOffsets = [], []
		off = 0
		for cell in range(self.layers * self.ndir):
			insz = self.layerInsize(cell // self.ndir)
			self.matOffsets.append((off, off + G * H * insz))
			off += G * H * (insz + H)
		for cell in range(self.layers * self.ndir):
			self.biasOffsets.append((off, off + G * H))
			off += 2 * G * H
		self.wsize = off

	# ------------------------------------------------------------------------------------------ parameter views
	def _view(self, W, offset, shape):
		size = int(np.prod(shape))
		itemsize = W.dtype.itemsize
		return GPUArray(shape, W.dtype, gpudata=W.gpudata[offset * itemsize:(offset + size) * itemsize])

	def layerInsize(self, layer):
		return self.insize if layer == 0 else self.ndir * self.hsize

	def stacked(self, W, cell):
		"""(Wcat, Rcat, bw, br) of one layer direction: the stacked gate matrices the GEMMs use."""
		H, G, insz = self.hsize, self.ngates, self.layerInsize(cell // self.ndir)
		woff, roff = self.matOffsets[cell]
		bwoff, broff = self.biasOffsets[cell]
		return (self._view(W, woff, (G * H, insz)), self._view(W, roff, (G * H, H)), self._view(W, bwoff, (G * H, )),
				self._view(W, broff, (G * H, )))

	def getParam(self, W, cell, linLayer):
		"""((Woffset, wsize), (biasOffset, biasSize)) in elements, as CuDnn.Rnn.getParam returns (Cuda/Backend.py:205-218);
		`cell` counts layer directions like the reference's `layer` argument does for bidirectional nets."""
		H, G, insz = self.hsize, self.ngates, self.layerInsize(cell // self.ndir)
		woff, roff = self.matOffsets[cell]
		bwoff, broff = self.biasOffsets[cell]
		if linLayer < G:
			return (woff + linLayer * H * insz, H * insz), (bwoff + linLayer * H, H)
		g = linLayer - G
		return (roff + g * H * H, H * H), (broff + g * H, H)

	def acquireParams(self, W):
		"""List (one dict per layer direction) of named views into W -- the reference's acquireRnnParams."""
		H, G = self.hsize, self.ngates
		params = []
		for cell in range(self.layers * self.ndir):
			insz = self.layerInsize(cell // self.ndir)
			cellparams = {}
			for linLayer in range(2 * G):
				wtype = "w" if linLayer < G else "r"
				gate = _GATES[self.mode][linLayer % G]
				(woff, wsize), (boff, bsize) = self.getParam(W, cell, linLayer)
				cellparams["%s%s" % (wtype, gate)] = self._view(W, woff, (H, insz if wtype == "w" else H))
				cellparams["b%s%s" % (wtype, gate)] = self._view(W, boff, (bsize, ))
			params.append(cellparams)
		return params

	# ------------------------------------------------------------------------------------------ one layer direction
	def _forwardCell(self, x, W, cell, reverse, h0, c0, allocator):
		T, B, insz = x.shape
		H, G = self.hsize, self.ngates
		blas = self.backend.blas
		Wcat, Rcat, bw, br = self.stacked(W, cell)

		# input projection of every step in one GEMM: (T*B, in) x (in, G*H)
		gates = GPUArray((T, B, G * H), _f32, allocator=allocator)
		blas.gemm(x.reshape(T * B, insz), Wcat, gates.reshape(T * B, G * H), transpB=True)

		plain = self.mode in (MODE_RELU, MODE_TANH)
		out = gates if plain else GPUArray((T, B, H), _f32, allocator=allocator)       # plain RNN: activations in place
		cellsbuf = GPUArray((T, B, H), _f32, allocator=allocator) if self.mode == MODE_LSTM else None
		rec = GPUArray((T, B, G * H), _f32, allocator=allocator) if self.mode == MODE_GRU else None

		order = range(T - 1, -1, -1) if reverse else range(T)
		prev = None
		for t in order:
			hprev = h0 if prev is None else out[prev]
			if self.mode == MODE_GRU:
				if hprev is not None:
					blas.gemm(hprev, Rcat, rec[t], transpB=True)
				else:
					check(lib.pz_memset8(rec[t].ptr, 0, rec[t].nbytes, None))
				check(lib.pz_gru_cell_fwd(gates[t].ptr, rec[t].ptr, bw.ptr, br.ptr, hprev.ptr if hprev is not None else None, out[t].ptr,
										  B, H, None))
			else:
				if hprev is not None:
					blas.gemm(hprev, Rcat, gates[t], transpB=True, alpha=1.0, beta=1.0)
				if self.mode == MODE_LSTM:
					cprev = c0 if prev is None else cellsbuf[prev]
					check(lib.pz_lstm_cell_fwd(gates[t].ptr, bw.ptr, br.ptr, cprev.ptr if cprev is not None else None, cellsbuf[t].ptr,
											   out[t].ptr, B, H, None))
				else:
					check(lib.pz_rnn_cell_fwd(gates[t].ptr, bw.ptr, br.ptr, B, H, self.mode, None))
			prev = t
		return out, {"acts": gates, "cells": cellsbuf, "rec": rec, "indata": x, "out": out, "h0": h0, "c0": c0, "reverse": reverse}

	def _backwardCell(self, dy, W, cell, saved, allocator):
		"""dy: (T, B, H) gradient of this direction's outputs -> (dx, dh0, dc0); stores the gate gradients in `saved`."""
		acts, cellsbuf, rec, out, h0, c0 = saved["acts"], saved["cells"], saved["rec"], saved["out"], saved["h0"], saved["c0"]
		T, B, H = out.shape
		G = self.ngates
		insz = saved["indata"].shape[2]
		blas = self.backend.blas
		Wcat, Rcat, _, _ = self.stacked(W, cell)

		dgates = GPUArray((T, B, G * H), _f32, allocator=allocator)
		drec = GPUArray((T, B, G * H), _f32, allocator=allocator) if self.mode == MODE_GRU else dgates
		dhnext = GPUArray((B, H), _f32, allocator=allocator)
		dc = GPUArray((B, H), _f32, allocator=allocator) if self.mode == MODE_LSTM else None

		fwdorder = list(range(T - 1, -1, -1) if saved["reverse"] else range(T))
		for pos in range(T - 1, -1, -1):
			t = fwdorder[pos]
			tprev = fwdorder[pos - 1] if pos > 0 else None
			last = pos == T - 1
			hprev = h0 if tprev is None else out[tprev]
			nxt = None if last else dhnext.ptr
			if self.mode == MODE_LSTM:
				cprev = c0 if tprev is None else cellsbuf[tprev]
				check(lib.pz_lstm_cell_bwd(dy[t].ptr, nxt, dc.ptr, acts[t].ptr, cellsbuf[t].ptr, cprev.ptr if cprev is not None else None,
										   dgates[t].ptr, B, H, 1 if last else 0, None))
				blas.gemm(dgates[t], Rcat, dhnext)
			elif self.mode == MODE_GRU:
				check(lib.pz_gru_cell_bwd(dy[t].ptr, nxt, acts[t].ptr, rec[t].ptr, hprev.ptr if hprev is not None else None, dgates[t].ptr,
										  drec[t].ptr, dhnext.ptr, B, H, None))
				blas.gemm(drec[t], Rcat, dhnext, alpha=1.0, beta=1.0)          # dh[t-1] = dh * i + drec R
			else:
				check(lib.pz_rnn_cell_bwd(dy[t].ptr, nxt, out[t].ptr, dgates[t].ptr, B * H, self.mode, None))
				blas.gemm(dgates[t], Rcat, dhnext)

		saved["dgates"], saved["drec"], saved["order"] = dgates, drec, fwdorder
		# input gradient of every step in one GEMM: (T*B, G*H) x (G*H, in)
		dx = GPUArray((T, B, insz), _f32, allocator=allocator)
		blas.gemm(dgates.reshape(T * B, G * H), Wcat, dx.reshape(T * B, insz))
		return dx, dhnext, dc

	def _paramsCell(self, dw, cell, saved, allocator):
		T, B, H = saved["out"].shape
		G = self.ngates
		blas, matmod = self.backend.blas, self.backend.matmod
		dWcat, dRcat, dbw, dbr = self.stacked(dw, cell)
		dgates, drec, x, out, h0 = saved["dgates"], saved["drec"], saved["indata"], saved["out"], saved["h0"]
		insz = x.shape[2]

		dg2 = dgates.reshape(T * B, G * H)
		blas.gemm(dg2, x.reshape(T * B, insz), dWcat, transpA=True)                                       # (G*H, in)
		# dRcat = sum_t drec[t]^T h[t-1]: in forward time order the previous step is t-1 (t+1 for the reversed direction)
		if T > 1:
			if saved["reverse"]:
				blas.gemm(drec[:T - 1].reshape((T - 1) * B, G * H), out[1:].reshape((T - 1) * B, H), dRcat, transpA=True)
			else:
				blas.gemm(drec[1:].reshape((T - 1) * B, G * H), out[:T - 1].reshape((T - 1) * B, H), dRcat, transpA=True)
		if h0 is not None:
			first = saved["order"][0]
			blas.gemm(drec[first], h0, dRcat, transpA=True, alpha=1.0, beta=1.0 if T > 1 else 0.0)
		matmod.matsum(dg2, axis=0, out=dbw)
		if self.mode == MODE_GRU:
			matmod.matsum(drec.reshape(T * B, G * H), axis=0, out=dbr)
		else:
			dbr.set(dbw)

	# ------------------------------------------------------------------------------------------ inter-layer dropout
	# cudnnSetRNNDescriptor's dropout (CuDnnRnn.c: the dropout descriptor built from `dropout` and `seed`): applied in training to
	# the output of every layer but the last, same kernel and keep rule as Modules/Dropout.py (keep when the random word is below
	# (1 - p) * UINT_MAX, scale by 1 / (1 - p)); the masks come from this descriptor's own generator, seeded with `seed`.  cuDNN's
	# random stream is not reproducible outside cuDNN -- the statistics are, and the backward pass reuses the stored words.
	def _dropout(self, x, reserve, test, allocator):
		if self.dropout == 0.0 or test:
			reserve.rands.append(None)
			return x
		if self.rng is None:
			from .backend import RandomNumberGenerator
			self.rng = RandomNumberGenerator(seed=int(self.seed))
		rands = GPUArray(x.shape, np.dtype(np.uint32), allocator=allocator)
		self.rng.fillInteger(rands)
		reserve.rands.append(rands)
		out = GPUArray(x.shape, _f32, allocator=allocator)
		self.backend.dropoutKer(_f32)(out, x, rands, self._partition(), 1.0 - self.dropout)      # the kernel takes the KEEP probability
		return out

	def _partition(self):
		return int((1.0 - self.dropout) * np.iinfo(np.uint32).max)

	def _dropoutBackward(self, dy, layer, reserve, allocator):
		rands = reserve.rands[layer] if layer < len(reserve.rands) else None
		if rands is None:
			return dy
		out = GPUArray(dy.shape, _f32, allocator=allocator)
		self.backend.dropoutKer(_f32)(out, dy, rands, self._partition(), 1.0 - self.dropout)
		return out

	# ------------------------------------------------------------------------------------------ public surface
	def forward(self, data, W, hidden=None, cells=None, test=False, allocator=None):
		if data.ndim != 3 or data.shape[2] != self.insize or data.dtype != _f32:
			raise ValueError("invalid rnn input layout %s" % (data.shape, ))
		if W.dtype != _f32 or W.size != self.wsize:
			raise ValueError("invalid rnn weights size")
		reserve = RnnReserve()

		x = data
		for layer in range(self.layers):
			outs = []
			for d in range(self.ndir):
				cell = layer * self.ndir + d
				h0 = None if hidden is None else hidden[cell]
				c0 = None if cells is None else cells[cell]
				out, saved = self._forwardCell(x, W, cell, d == 1, h0, c0, allocator)
				outs.append(out)
				reserve.cells.append(saved)
			x = outs[0] if self.ndir == 1 else self.backend.concatenate(outs, 2, None, allocator=allocator)
			reserve.outs.append(x)
			if layer < self.layers - 1:
				x = self._dropout(x, reserve, test, allocator)
		return x if test else (x, reserve)

	def backwardData(self, grad, outdata, W, reserve, hidden=None, cells=None, allocator=None):
		if grad.shape != outdata.shape:
			raise ValueError("invalid rnn gradient layout %s" % (grad.shape, ))
		H = self.hsize
		ncells = self.layers * self.ndir
		dhx, dcx = [None] * ncells, [None] * ncells

		dy = grad
		for layer in range(self.layers - 1, -1, -1):
			if layer < self.layers - 1:
				dy = self._dropoutBackward(dy, layer, reserve, allocator)          # the mask layer `layer`'s output went through
			parts = [dy] if self.ndir == 1 else self.backend.split(dy, (H, H), 2, allocator=allocator)
			dx = None
			for d in range(self.ndir):
				cell = layer * self.ndir + d
				dxd, dhx[cell], dcx[cell] = self._backwardCell(parts[d], W, cell, reserve.cells[cell], allocator)
				if dx is None:
					dx = dxd
				else:
					self.backend.toVectorAddVectorKer(_f32)(dx, dxd, 1.0)
			dy = dx
		reserve.grads = True
		return dy, dhx, dcx

	def backwardParams(self, data, outdata, reserve, hidden=None, allocator=None):
		if not reserve.grads:
			raise ValueError("backwardParams needs the reserve of a backwardData call")
		dw = GPUArray.zeros((self.wsize, ), _f32, allocator=allocator)
		for cell in range(self.layers * self.ndir):
			self._paramsCell(dw, cell, reserve.cells[cell], allocator)
		return dw
