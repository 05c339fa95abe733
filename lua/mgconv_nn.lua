--[[ mgconv_nn.lua -- Torch7 nn.Module classes on top of libmgconv (LuaJIT FFI).

   nn.MGConvBN is the drop-in for the per-scale chain the reference's builders assemble
   (models/ilsvrc/rnmg.lua:53-82 + 22-39):

       ConcatTable{ Seq{SelectTable(i-1), SpatialMaxPooling(2,2,2,2):ceil()}, SelectTable(i),
                    Seq{SelectTable(i+1), SpatialUpSamplingNearest(2)} } -> JoinTable(2)
         -> cudnn.SpatialConvolution(nIP, nOP, k,k, 1,1, p,p) -> nn.SpatialBatchNormalization(nOP, eps) [-> ReLU]

   input  : table {finer or nil, same, coarser or nil} of mgconv grids (NHWC bf16, see M.import)
   output : one grid (post BN/ReLU) plus its pooled companion for the next coarser scale.
   The call sequence is the one engine.py / ops.py replay from Python (ConvOp.fwd, ApplyOp.fwd,
   ApplyOp.bwd, ConvOp.bwd); parameters keep Torch's layout (weight [nOP][nIP][k][k], input planes
   ordered finer | same | coarser) so getParameters(), optim.sgd and torch.save work unchanged.
   Raw device pointers are fetched on every call (getParameters() re-homes the tensors,
   pipelines/standard/train.lua:115); nothing the shim owns is serialised: clearState() drops it.

   UNTESTED HERE: no LuaJIT/Torch7 exists in the build container or on the GPU box.
]]
local mg = require 'mgconv_ffi'
local ffi, C = mg.ffi, mg.C

local MGConvBN, parent = torch.class('nn.MGConvBN', 'nn.Module')

function MGConvBN:__init(nIPs, nOP, k, eps, relu)
   parent.__init(self)
   self.nIPs, self.nOutputPlane, self.kW, self.kH = nIPs, nOP, k, k   -- nIPs = {finer, same, coarser} (0 = absent)
   self.nInputPlane = nIPs[1] + nIPs[2] + nIPs[3]
   self.eps, self.momentum, self.relu = eps or 1e-5, 0.1, relu and 1 or 0
   self.weight = torch.Tensor(nOP, self.nInputPlane, k, k)
   self.bias = torch.Tensor(nOP)
   self.gradWeight = torch.Tensor(nOP, self.nInputPlane, k, k):zero()
   self.gradBias = torch.Tensor(nOP):zero()
   self.bn_weight, self.bn_bias = torch.Tensor(nOP):fill(1), torch.Tensor(nOP):zero()
   self.bn_gradWeight, self.bn_gradBias = torch.Tensor(nOP):zero(), torch.Tensor(nOP):zero()
   self.running_mean, self.running_var = torch.zeros(nOP), torch.ones(nOP)
   self.train = true
   self:reset()
end

function MGConvBN:reset()   -- ConvInit of models/ilsvrc/rnmg.lua:288-294 (a module of a new type is invisible to findModules)
   local n = self.kW * self.kH * self.nOutputPlane
   self.weight:normal(0, math.sqrt(2 / n)); self.bias:zero()
end

function MGConvBN:parameters()
   return {self.weight, self.bias, self.bn_weight, self.bn_bias},
          {self.gradWeight, self.gradBias, self.bn_gradWeight, self.bn_gradBias}
end

local function dptr(t) return ffi.cast('void*', torch.pointer(t:storage())) end   -- cutorch: t:data()

function MGConvBN:_desc(input)
   local d = ffi.new('mg_conv_desc')
   local modes = {C.MG_SEG_SAME, C.MG_SEG_SAME, C.MG_SEG_UP}   -- finer enters as its pooled companion
   local n = 0
   for i = 1, 3 do
      if input[i] then d.seg[n] = input[i].grid; d.seg_mode[n] = modes[i]; n = n + 1 end
   end
   d.n_seg, d.ksize, d.stride, d.pad, d.Cout = n, self.kW, 1, (self.kW == 1) and 0 or 1, self.nOutputPlane
   d.H, d.W = input[2].grid.H, input[2].grid.W
   return d
end

function MGConvBN:updateOutput(input)
   local ctx, s = self.ctx, self.state
   local d = self:_desc(input)
   mg.check(ctx, C.mg_conv_pack_weights(ctx, d, self.weight:data(), s.wpack, 0))
   mg.check(ctx, C.mg_memset_zero(ctx, s.sums, 16 * self.nOutputPlane))
   mg.check(ctx, C.mg_conv_forward(ctx, d, self.weight:data(), s.wpack, self.bias:data(), s.y, s.sums))
   mg.check(ctx, C.mg_bn_finalize(ctx, s.sums, s.count, self.nOutputPlane, s.y.Cp, self.bn_weight:data(), self.bn_bias:data(),
                                  self.running_mean:data(), self.running_var:data(), self.eps, self.momentum,
                                  self.train and 1 or 0, s.scale, s.shift, s.mean, s.invstd))
   s.z = mg.grid(s.y.data, s.y.N, s.y.H, s.y.W, s.y.C, s.scale, s.shift, 0)
   mg.check(ctx, C.mg_residual_forward(ctx, s.z, self.shortcut and self.shortcut.grid or nil, self.relu, s.out, s.pooled))
   self.output = {grid = s.out, pooled = s.pooled}
   return self.output
end

function MGConvBN:backward(input, gradSources, scale)
   -- gradSources: array of mg_grad_src registered by the consumers of self.output (ConcatTable's
   -- backward-sum in gather form, see ops.py:Combine); returns dcat, whose slices the producers read
   local ctx, s = self.ctx, self.state
   local d = self:_desc(input)
   mg.check(ctx, C.mg_memset_zero(ctx, s.dsums, 16 * self.nOutputPlane))
   mg.check(ctx, C.mg_grad_combine(ctx, s.out, self.relu, s.y, #gradSources, s.srcs(gradSources), s.D, s.dsums))
   mg.check(ctx, C.mg_bn_backward(ctx, s.y, s.D, s.G, s.dsums, s.count, self.bn_weight:data(), s.mean, s.invstd,
                                  self.bn_gradWeight:data(), self.bn_gradBias:data(), scale or 1, s.coef,
                                  self.gradBias:data()))   -- conv gradBias fused into the BN-backward pass
   mg.check(ctx, C.mg_conv_backward_weight(ctx, d, s.G, self.gradWeight:data(), nil, scale or 1))
   mg.check(ctx, C.mg_conv_pack_weights(ctx, d, self.weight:data(), s.wpack_t, 1))
   mg.check(ctx, C.mg_conv_backward_data(ctx, d, self.weight:data(), s.wpack_t, s.G, s.dcat))
   self.gradInput = s.dcat
   return self.gradInput
end

function MGConvBN:clearState()
   self.state, self.ctx = nil, nil   -- plans / workspaces are rebuilt lazily; never serialised
   return parent.clearState(self)
end
