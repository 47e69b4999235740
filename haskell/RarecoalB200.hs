{-# LANGUAGE ForeignFunctionInterface #-}
-- | Binding of librarecoal_b200.so (include/rarecoal_b200.h) for rarecoal's likelihood path.
--
-- Drop-in for the per-pattern fan-out of RarecoalLib.MaxUtils.computeFrequencySpectrum
-- (src/RarecoalLib/MaxUtils.hs:107-110): instead of
--
-- >   patternProbs <- sequence $ parMap rdeepseq (C1.getProb modelSpec jss nVec) patterns
--
-- one `rc_eval` call returns the probability of every pattern (and the log-likelihood) from the GPU.
-- This module was written against the reference at v1.5.1; it cannot be compiled in the build image of this
-- repository (no ghc) and is shipped as source for the maintainer.  See INTEGRATION.md.
module RarecoalLib.RarecoalB200
    ( B200Context, withB200, withB200Multi, setProblem, setTimes, evalModels, B200Result(..)
    ) where

import           Control.Exception     (bracket, throwIO)
import           Control.Monad         (when)
import           Data.Int              (Int32, Int64)
import qualified Data.Vector.Storable  as VS
import           Foreign
import           Foreign.C.String      (CString, peekCString)
import           Foreign.C.Types

import           RarecoalLib.StateSpace (EventType (..), ModelEvent (..), ModelSpec (..))
import           RarecoalLib.Utils      (RarecoalException (..))

data RcCtx
newtype B200Context = B200Context (Ptr RcCtx)

-- | rc_event: { double t; int32 type; int32 k; int32 l; double v }  (32 bytes: v sits at offset 24 after padding)
data RcEvent = RcEvent !Double !Int32 !Int32 !Int32 !Double

instance Storable RcEvent where
    sizeOf _ = 32
    alignment _ = 8
    peek p = RcEvent <$> peekByteOff p 0 <*> peekByteOff p 8 <*> peekByteOff p 12 <*> peekByteOff p 16
                     <*> peekByteOff p 24
    poke p (RcEvent t ty k l v) =
        pokeByteOff p 0 t >> pokeByteOff p 8 ty >> pokeByteOff p 12 k >> pokeByteOff p 16 l >> pokeByteOff p 24 v

toRcEvent :: ModelEvent -> RcEvent
toRcEvent (ModelEvent t e) = case e of
    Join k l       -> RcEvent t 0 (fi k) (fi l) 0
    Split k l m    -> RcEvent t 1 (fi k) (fi l) m
    SetPopSize k p -> RcEvent t 2 (fi k) 0 p
    SetFreeze k b  -> RcEvent t 3 (fi k) 0 (if b then 1 else 0)
  where fi = fromIntegral

-- `safe` calls from a bound thread: rc_eval blocks on the GPU; one context per OS thread.
foreign import ccall safe "rc_create"      c_rc_create      :: CInt -> Ptr (Ptr RcCtx) -> IO CInt
-- one context that drives the first n GPUs of the process (devices = nullPtr): rc_eval then shards the models of a batch over
-- them, or the patterns of a single likelihood with an NCCL all-reduce of the partial sums — what parMap does with cores
foreign import ccall safe "rc_create_multi" c_rc_create_multi :: CInt -> Ptr Int32 -> Ptr (Ptr RcCtx) -> IO CInt
foreign import ccall safe "rc_destroy"     c_rc_destroy     :: Ptr RcCtx -> IO ()
foreign import ccall safe "rc_last_error"  c_rc_last_error  :: Ptr RcCtx -> IO CString
foreign import ccall safe "rc_set_problem" c_rc_set_problem ::
    Ptr RcCtx -> CInt -> CInt -> CInt -> Ptr Int32 -> Ptr Int32 -> Ptr Int64 -> Int64 -> IO CInt
foreign import ccall safe "rc_set_times"   c_rc_set_times   :: Ptr RcCtx -> CInt -> Ptr Double -> IO CInt
-- mNoShortcut is a field of every ModelSpec (StateSpace.hs:48): rc_eval_models takes one flag per model
foreign import ccall safe "rc_eval_models" c_rc_eval_models ::
    Ptr RcCtx -> CInt -> Ptr RcEvent -> Ptr Int32 -> Ptr Double -> Ptr Double -> Ptr Int32 -> Double
    -> Ptr Double -> Ptr Double -> Ptr Int32 -> IO CInt

check :: Ptr RcCtx -> CInt -> IO ()
check ctx rc = when (rc /= 0) $ do
    msg <- c_rc_last_error ctx >>= peekCString
    throwIO $ case rc of
        1 -> RarecoalModelException msg
        2 -> RarecoalCompException msg
        _ -> RarecoalCompException ("rarecoal_b200 (" ++ show rc ++ "): " ++ msg)

-- | One context per GPU; there is no CPU fallback — creation fails without a CUDA device.
withB200 :: Int -> (B200Context -> IO a) -> IO a
withB200 device = bracket create destroy
  where
    create = alloca $ \pp -> do
        rc <- c_rc_create (fromIntegral device) pp
        when (rc /= 0) $ throwIO (RarecoalCompException "rc_create failed: no usable CUDA device")
        B200Context <$> peek pp
    destroy (B200Context p) = c_rc_destroy p

-- | The same over the first @n@ GPUs of the machine (@--nrGpus@): every call below works unchanged on such a context.
withB200Multi :: Int -> (B200Context -> IO a) -> IO a
withB200Multi n = bracket create destroy
  where
    create = alloca $ \pp -> do
        rc <- c_rc_create_multi (fromIntegral n) nullPtr pp
        when (rc /= 0) $ throwIO (RarecoalCompException ("rc_create_multi failed for " ++ show n ++ " devices"))
        B200Context <$> peek pp
    destroy (B200Context p) = c_rc_destroy p

-- | What computeFrequencySpectrum closes over (MaxUtils.hs:97-107): standard-order patterns and nVec already
-- mapped to model-branch order, counts (0 for unseen patterns) and raTotalNrSites.
setProblem :: B200Context -> Int -> Int -> [[Int]] -> [Int] -> [Int64] -> Int64 -> IO ()
setProblem (B200Context ctx) nrPops maxAf patterns nVec counts total =
    VS.unsafeWith (VS.fromList (map fromIntegral (concat patterns))) $ \pp ->
    VS.unsafeWith (VS.fromList (map fromIntegral nVec)) $ \pn ->
    VS.unsafeWith (VS.fromList counts) $ \pc ->
        c_rc_set_problem ctx (fromIntegral nrPops) (fromIntegral maxAf) (fromIntegral (length patterns)) pp pn pc total
            >>= check ctx

setTimes :: B200Context -> [Double] -> IO ()
setTimes (B200Context ctx) ts =
    VS.unsafeWith (VS.fromList ts) $ \pt -> c_rc_set_times ctx (fromIntegral (length ts)) pt >>= check ctx

data B200Result = B200Result
    { brProbs  :: VS.Vector Double   -- ^ getProb of every pattern, standard order
    , brLogl   :: Double             -- ^ computeLogLikelihoodFromSpec
    , brStatus :: Int                -- ^ 0, 1 = RarecoalModelException, 2 = RarecoalCompException
    }

-- | A batch of ModelSpecs sharing the time grid (logl: one; maxl / mcmc / find: many).
evalModels :: B200Context -> Int -> Double -> [ModelSpec] -> IO [B200Result]
evalModels (B200Context ctx) nPat siteRed specs = do
    let b      = length specs
        evs    = concatMap (map toRcEvent . mEvents) specs
        offs   = scanl (+) 0 (map (length . mEvents) specs)
        thetas = map mTheta specs
        discs  = concatMap mDiscoveryRates specs
        noScs  = map (\m -> if mNoShortcut m then 1 else 0) specs :: [Int32]
    VS.unsafeWith (VS.fromList evs) $ \pev ->
      VS.unsafeWith (VS.fromList (map fromIntegral offs :: [Int32])) $ \poff ->
      VS.unsafeWith (VS.fromList thetas) $ \pth ->
      VS.unsafeWith (VS.fromList discs) $ \pdi ->
      VS.unsafeWith (VS.fromList noScs) $ \pns ->
      allocaArray (b * nPat) $ \pprobs ->
      allocaArray b $ \plogl ->
      allocaArray b $ \pstat -> do
        c_rc_eval_models ctx (fromIntegral b) pev poff pth pdi pns siteRed pprobs plogl pstat >>= check ctx
        probs <- VS.fromList <$> peekArray (b * nPat) pprobs
        logls <- peekArray b plogl
        stats <- peekArray b pstat
        return [ B200Result (VS.slice (i * nPat) nPat probs) l (fromIntegral s)
               | (i, l, s) <- zip3 [0 ..] logls (stats :: [Int32]) ]
