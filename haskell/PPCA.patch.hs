-- UNVERIFIED: replacement bodies for src/GPLVM/PPCA.hs (see INTEGRATION.md section 3)
emStepsFast learnMatrix initMatrix initVariance stopParam =          -- was PPCA.hs:60-94
  let (w, var, ll) = ppcaDense (computeS learnMatrix) (computeS initMatrix) initVariance stopParam
  in (delay w, var, ll, Nothing)

emStepsMissed learnMatrix initMatrix initVariance stopParam =        -- was PPCA.hs:102-182
  let (w, var, ll, restored) = ppcaMissing (computeS learnMatrix) (computeS initMatrix) initVariance stopParam
  in (delay w, var, ll, Just (delay restored))
