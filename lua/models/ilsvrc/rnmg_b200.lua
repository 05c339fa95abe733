--[[ models/ilsvrc/rnmg_b200.lua -- R-MG-18/34 with the residual multigrid units on libmgconv (B200).

   Drop this file next to models/ilsvrc/rnmg.lua of the reference and select it with `-netType ilsvrc/rnmg_b200`
   (model.lua:23 loads models/<netType>.lua).  It IS the reference's model file with one binding changed: the local
   `mgConv` builds nn.MGStage (lua/mgconv_nn.lua) instead of the Concat/Max/UpSample/JoinTable/Convolution/BatchNorm graph
   (rnmg.lua:91-159).  Stem (161-189), mgPool (191-224), classifier (280-286), criterion (325-329), training rule and the
   NET hooks are the reference's own: `NET` starts as a copy of what models/ilsvrc/rnmg.lua returns.
]]
local ref = paths.dofile('rnmg.lua')          -- the reference's NET table: trainRule, trainOutput, testOutput, arguments ...
local mgb = require 'mgconv_nn'

local NET = {}
for k, v in pairs(ref) do NET[k] = v end

local Convolution = cudnn.SpatialConvolution
local Avg = cudnn.SpatialAveragePooling
local Max = nn.SpatialMaxPooling
local ReLU = nn.ReLU
local SBatchNorm = nn.SpatialBatchNormalization

-- rnmg.lua:161-189, unchanged (the 7x7 stem runs on stock cudnn modules; libmgconv has its own stem kernel for hosts that
-- drive the whole network through the C ABI)
local function mgConvInput(nOutputPlanes)
   local resample_image = nn.ConcatTable()
   for iG = 1, #nOutputPlanes do
      local proc = nn.Sequential()
      if iG > 1 then
         local r = torch.pow(2, iG - 1)
         proc:add(Avg(r, r, r, r, 0, 0))
      end
      proc:add(Convolution(3, nOutputPlanes[iG], 7, 7, 2, 2, 3, 3))
      proc:add(SBatchNorm(nOutputPlanes[iG]))
      proc:add(ReLU(true))
      proc:add(Max(3, 3, 2, 2, 1, 1))
      resample_image:add(proc)
   end
   return resample_image
end

-- rnmg.lua:191-224, unchanged: per-grid MaxPool 2x2 s2 ceil; isConcat joins the pooled grid n-1 with grid n.
-- Mutates nInputPlanes in place exactly like the reference (210-211).
local function mgPool(nInputPlanes, isConcat)
   local nGrids = #nInputPlanes
   local pool = nn.ConcatTable()
   if isConcat then
      for iG = 1, nGrids - 2 do
         pool:add(nn.Sequential():add(nn.SelectTable(iG)):add(Max(2, 2, 2, 2, 0, 0):ceil()))
      end
      local last = nn.Sequential()
      local both = nn.ConcatTable()
      both:add(nn.Sequential():add(nn.SelectTable(nGrids - 1)):add(Max(2, 2, 2, 2, 0, 0):ceil()))
      both:add(nn.SelectTable(nGrids))
      last:add(both):add(nn.JoinTable(2))
      pool:add(last)
      nInputPlanes[nGrids - 1] = nInputPlanes[nGrids - 1] + nInputPlanes[nGrids]
      nInputPlanes[nGrids] = nil
   else
      for iG = 1, nGrids do
         pool:add(nn.Sequential():add(nn.SelectTable(iG)):add(Max(2, 2, 2, 2, 0, 0):ceil()))
      end
   end
   return pool
end

function NET.createModel(opt)
   NET.packages()
   local model = nn.Sequential()
   local inputBlock = {64, 32, 16}                       -- (224,112,56) -> (56,28,14)
   model:add(mgConvInput(inputBlock))
   local cfg = { [18] = {2, 2, 2, 2}, [34] = {3, 4, 6, 3} }
   local blocks = {
      {{64, 32, 16}, {3, 3, 3}, false},                  -- (56,28,14) -> (28,14,7)
      {{128, 64, 32}, {3, 3, 3}, true},                  -- (28,14,7)  -> (14,7)
      {{256, 128}, {3, 3}, true},                        -- (14,7)     -> (7)
      {{512}, {3}, false},
   }
   local nIPs = inputBlock
   for indBlock = 1, #blocks do
      local nOPs, kernels, isConcat = blocks[indBlock][1], blocks[indBlock][2], blocks[indBlock][3]
      for indLayer = 1, cfg[opt.depth][indBlock] do
         local ips = {}
         for i, c in ipairs(nIPs) do ips[i] = c end
         model:add(mgb.mgConv(ips, nOPs, kernels))       -- <- the one changed line: nn.MGStage instead of the module graph
         nIPs = {}
         for i, depth in ipairs(nOPs) do nIPs[i] = depth end
      end
      if indBlock < #blocks then model:add(mgPool(nIPs, isConcat)) end
   end
   local classifier = nn.Sequential()
   classifier:add(nn.SelectTable(1))
   classifier:add(Avg(7, 7, 1, 1, 0, 0))
   classifier:add(nn.View(-1, nIPs[1]))
   classifier:add(nn.Linear(nIPs[1], 1000))
   classifier:add(nn.LogSoftMax())
   model:add(classifier)

   -- ConvInit / BNInit (rnmg.lua:288-308) for the stock modules of the stem; nn.MGStage initialised itself the same way
   for _, name in ipairs{'cudnn.SpatialConvolution', 'nn.SpatialConvolution'} do
      for _, v in pairs(model:findModules(name)) do
         v.weight:normal(0, math.sqrt(2 / (v.kW * v.kH * v.nOutputPlane))); v.bias:zero()
      end
   end
   for _, name in ipairs{'cudnn.SpatialBatchNormalization', 'nn.SpatialBatchNormalization'} do
      for _, v in pairs(model:findModules(name)) do v.weight:fill(1); v.bias:zero() end
   end
   for _, v in pairs(model:findModules('nn.Linear')) do v.bias:zero() end
   model:get(1).gradInput = nil
   if opt.nGPU > 1 then
      return makeDataParallel(model, opt.nGPU, NET)      -- multigpu.lua:81-103; one mg_ctx per GPU / Lua state (mgconv_nn.lua)
   end
   return model
end

function NET.arguments(cmd)
   if ref.arguments then ref.arguments(cmd) end
   cmd:option('-precision', 'bf16', 'bf16 (tcgen05 tensor cores) | fp32 (CUDA cores, parity mode)')
end

return NET
